// cusift_b200 — reference-compatible image container.
// Same public layout and methods as danielsuo/cuSIFT cuImage.h:8-26 (48-byte
// object: width, height, pitch [floats], h_data, d_data, t_data, two ownership
// flags); `CudaImage` is the upstream CudaSift name of the same class.
#ifndef CUSIFT_B200_CUIMAGE_H
#define CUSIFT_B200_CUIMAGE_H

#include <cstddef>

class cuImage {
public:
  int width, height;
  int pitch;              // row stride of d_data in floats
  float *h_data;          // host pixels (borrowed unless h_internalAlloc)
  float *d_data;          // device pixels (borrowed unless d_internalAlloc)
  float *t_data;          // unused (kept for layout compatibility)
  bool d_internalAlloc;
  bool h_internalAlloc;

public:
  cuImage();
  // Wraps dense host pixels (row stride = width), allocates the device image with
  // pitch iAlignUp(width,128) and, when `download` is set, uploads it.
  cuImage(int width, int height, float *h_data, bool download = true);
  ~cuImage();

  void AllocateWithHostMemory(int width, int height, float *h_data);
  void Allocate(int width, int height, int pitch, bool withHost, float *d_data = NULL, float *h_data = NULL);
  double DeviceToHost();   // returns elapsed ms
  double HostToDevice();   // returns elapsed ms
};

typedef cuImage CudaImage;

#endif
