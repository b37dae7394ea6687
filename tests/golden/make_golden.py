"""Regenerates tests/golden/ from the reference's own fixtures (run in the build
container, where /root/reference is mounted; the GPU box only sees the outputs).

Copies DATA only (no reference source code):
  sift1.bin, sift2.bin          test/data/sift/sift{1,2}       VLFeat descriptors (test/test.cpp:27-28)
  match_indices1_2.bin          test/data/match_indices/...    326 MATLAB NN pairs (test/test.cpp:30-40)
  cusift1_check.bin             test/data/cusift1_check        4096 x {x,y,scale,orientation} (test/detector.cpp:65-84)
  match1_2.bin                  test/data/match/match1_2       326 x (xyz, xyz) doubles (test/test.cpp:42; debug.cpp:244-279)
  rigid_ransac.bin              test/data/RigidTransform_RANSAC.bin  120 3-D matches, 10 x 3 indices, MATLAB Rt
                                (test/test.cpp:58-110; format: extras/debug.cpp:318-372)
  frames.npz                    gray1 = test/data/gray1 (== cv2.imread(color1.jpg, 0), verified below)
                                gray2 = cv2.imread(test/data/color2.jpg, 0)      stored as uint8
"""
import shutil
import sys
from pathlib import Path

import cv2
import numpy as np

REF = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference") / "test" / "data"
OUT = Path(__file__).resolve().parent

shutil.copyfile(REF / "sift" / "sift1", OUT / "sift1.bin")
shutil.copyfile(REF / "sift" / "sift2", OUT / "sift2.bin")
shutil.copyfile(REF / "match_indices" / "match_indices1_2", OUT / "match_indices1_2.bin")
shutil.copyfile(REF / "match" / "match1_2", OUT / "match1_2.bin")
shutil.copyfile(REF / "cusift1_check", OUT / "cusift1_check.bin")
shutil.copyfile(REF / "RigidTransform_RANSAC.bin", OUT / "rigid_ransac.bin")

gray1 = np.fromfile(REF / "gray1", np.float32).reshape(480, 640)
dec1 = cv2.imread(str(REF / "color1.jpg"), 0)
assert np.array_equal(gray1, dec1.astype(np.float32)), "gray1 must equal the decoded color1.jpg"
gray2 = cv2.imread(str(REF / "color2.jpg"), 0)
assert gray1.max() <= 255 and np.array_equal(gray1, np.round(gray1))
np.savez_compressed(OUT / "frames.npz", gray1=gray1.astype(np.uint8), gray2=gray2)
print("golden fixtures written to", OUT)
