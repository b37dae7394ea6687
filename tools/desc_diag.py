#!/usr/bin/env python
"""Development aid: where do our descriptors differ from the reference's?"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import cusift_b200 as csb  # noqa: E402
import parity_utils as PU  # noqa: E402
from oracle import oracle as O  # noqa: E402

out = ROOT / "gpurun_out"; out.mkdir(exist_ok=True)
work = out / "work"; work.mkdir(exist_ok=True)
g1, _ = PU.golden_frames()
ctx = csb.Context(0, 1)
ours = ctx.extract(g1, csb.make_params(6, 0.0, 0.1), max_pts=32768)
ours2 = ctx.extract(g1, csb.make_params(6, 0.0, 0.1), max_pts=32768)
ref = O.ref_extract(g1, work, 6, 0.0, 0.1, 10.0, 0.0, False, 32768, safe=True, tag="dd")
ref2 = O.ref_extract(g1, work, 6, 0.0, 0.1, 10.0, 0.0, False, 32768, safe=True, tag="dd2")
rep = {}
for name, a, b in (("ours_vs_ref", ours, ref), ("ours_vs_ours", ours, ours2), ("ref_vs_ref", ref, ref2)):
    ia, ib, _, _ = PU.match_sets(a, b)
    A, B = a[ia], b[ib]
    dr = PU.desc_rel_l2(A["data"], B["data"])
    do = PU.ang_diff_deg(A["orientation"], B["orientation"])
    rep[name] = {"n": len(ia), "desc_q": np.quantile(dr, [0.5, 0.9, 0.99, 1.0]).tolist(),
                 "ori_q": np.quantile(do, [0.5, 0.9, 0.99, 1.0]).tolist(),
                 "frac_desc_gt_1e-4": float((dr > 1e-4).mean()),
                 "corr_ori_desc": float(np.corrcoef(do, dr)[0, 1]) if len(ia) > 2 else None}
    if name == "ours_vs_ref":
        worst = np.argsort(-dr)[:6]
        rep["worst"] = []
        for wi in worst:
            d = (A["data"][wi] - B["data"][wi]).astype(float)
            top = np.argsort(-np.abs(d))[:10]
            rep["worst"].append({"rel": float(dr[wi]), "dori": float(do[wi]), "ori": float(B["orientation"][wi]),
                                 "scale": float(B["scale"][wi]), "sub": float(B["subsampling"][wi]),
                                 "bins": top.tolist(), "diffs": d[top].tolist(),
                                 "vals_ref": B["data"][wi][top].astype(float).tolist()})
        # bin-position statistics of the absolute differences
        D = np.abs(A["data"].astype(float) - B["data"].astype(float)).mean(0)
        rep["mean_abs_diff_by_cell"] = D.reshape(16, 8).sum(1).tolist()
        rep["mean_abs_diff_by_ori"] = D.reshape(16, 8).sum(0).tolist()
print(json.dumps(rep, indent=1))
(out / "desc_diag.json").write_text(json.dumps(rep, indent=1))
import shutil; shutil.rmtree(work, ignore_errors=True)
