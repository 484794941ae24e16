// Per-device "done once" flag for cudaFuncSetAttribute and device queries: those are per device, and a process may
// drive the library on more than one GPU (a benign race between threads only repeats a cheap call).
#pragma once
#include <cuda_runtime.h>

namespace gsmvi {

struct PerDeviceOnce {
  bool done[64] = {};
  static int dev() {
    int d = 0;
    cudaGetDevice(&d);
    return d;
  }
  bool get() const {
    const int d = dev();
    return d >= 0 && d < 64 && done[d];
  }
  void set() {
    const int d = dev();
    if (d >= 0 && d < 64) done[d] = true;
  }
};

struct PerDeviceInt {
  int v[64] = {};
  int& ref() {
    int d = PerDeviceOnce::dev();
    return v[(d >= 0 && d < 64) ? d : 0];
  }
};

}  // namespace gsmvi
