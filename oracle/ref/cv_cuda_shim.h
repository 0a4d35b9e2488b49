/* cv_cuda_shim.h -- the two OpenCV types the reference's PatchMatch kernels use, nothing else.
 *
 * TEST INFRASTRUCTURE (see oracle/ref/README.md). The reference's four __global__ kernels and three
 * __device__ functions (/root/reference/src/vehicle/patchmatch_gpu/patchmatch_gpu.cu:18-295) depend on
 * OpenCV only through cv::cuda::PtrStepSz<T>: a POD {data, step, rows, cols} whose operator()(y, x)
 * is `((T*)((char*)data + y * step))[x]` (opencv2/core/cuda_types.hpp, OpenCV 3.4: DevPtr / PtrStep /
 * PtrStepSz). The argument types of operator() are int, so float arguments -- MaskOcclusions passes
 * floats, patchmatch_gpu.cu:288-293 -- convert by truncation exactly as with the real header.
 * OpenCV is not installed in this image; this restates that public interface.
 */
#pragma once
#include <cassert>
#include <cstddef>
#include <cuda_runtime.h>

namespace cv { namespace cuda {

template <typename T> struct DevPtr {
  typedef T elem_type;
  typedef int index_type;
  enum { elem_size = sizeof(elem_type) };
  T* data;
  __host__ __device__ __forceinline__ DevPtr() : data(0) {}
  __host__ __device__ __forceinline__ DevPtr(T* data_) : data(data_) {}
  __host__ __device__ __forceinline__ size_t elemSize() const { return elem_size; }
  __host__ __device__ __forceinline__ operator T*() { return data; }
  __host__ __device__ __forceinline__ operator const T*() const { return data; }
};

template <typename T> struct PtrStep : public DevPtr<T> {
  __host__ __device__ __forceinline__ PtrStep() : step(0) {}
  __host__ __device__ __forceinline__ PtrStep(T* data_, size_t step_) : DevPtr<T>(data_), step(step_) {}
  size_t step;  /* stride between two consecutive rows in BYTES */
  __host__ __device__ __forceinline__ T* ptr(int y = 0) { return (T*)((char*)DevPtr<T>::data + y * step); }
  __host__ __device__ __forceinline__ const T* ptr(int y = 0) const { return (const T*)((const char*)DevPtr<T>::data + y * step); }
  __host__ __device__ __forceinline__ T& operator()(int y, int x) { return ptr(y)[x]; }
  __host__ __device__ __forceinline__ const T& operator()(int y, int x) const { return ptr(y)[x]; }
};

template <typename T> struct PtrStepSz : public PtrStep<T> {
  __host__ __device__ __forceinline__ PtrStepSz() : cols(0), rows(0) {}
  __host__ __device__ __forceinline__ PtrStepSz(int rows_, int cols_, T* data_, size_t step_)
      : PtrStep<T>(data_, step_), cols(cols_), rows(rows_) {}
  int cols;
  int rows;
};

namespace device {
__host__ __device__ __forceinline__ int divUp(int total, int grain) { return (total + grain - 1) / grain; }
}  // namespace device

}}  // namespace cv::cuda
