/* TEST INFRASTRUCTURE (oracle/): compatibility shim used ONLY to compile the UNMODIFIED reference sources
 * (/root/reference/source/reconstructionGPU2/{reconstruction_cuda2.cu,GPUWorker.cpp,GPUGauss/gaussfilter.cu})
 * into oracle/_ref/libref_cuda2.so with CUDA 12.9 for sm_100a.  It stands in for three things the 2015 sources
 * expect and this image does not have:
 *   (1) `helper_cuda.h` / `helper_functions.h` of the CUDA samples (checkCudaErrors, getLastCudaError);
 *   (2) the legacy texture-reference API removed in CUDA 12.0 (`texture<T,dim,mode>` file-scope references,
 *       `cudaBindTextureToArray`, `cudaBindTexture`, `tex3D(ref,..)`, `tex1Dfetch(ref,..)`): re-expressed on
 *       texture OBJECTS with the same address/filter/normalisation settings, so the hardware sampling path the
 *       reference relies on (linear filtering, border addressing) is the real one;
 *   (3) thrust API drift: `make_tuple<T...>(lvalue...)` with explicit template arguments (thrust 1.8 took
 *       const T&, libcu++ takes T&&) and unqualified `reduce` / `transform_reduce` calls that are ambiguous
 *       with cuda::std:: through ADL today.
 * Nothing here is arithmetic of the path.  No reference code is copied: this file only declares replacements
 * for external headers.  Not part of the product; never linked into libsvr_b200.so. */
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include <thrust/version.h>
#include <thrust/host_vector.h>
#include <thrust/device_vector.h>
#include <thrust/device_ptr.h>
#include <thrust/execution_policy.h>
#include <thrust/generate.h>
#include <thrust/sort.h>
#include <thrust/copy.h>
#include <thrust/fill.h>
#include <thrust/reduce.h>
#include <thrust/transform.h>
#include <thrust/transform_reduce.h>
#include <thrust/inner_product.h>
#include <thrust/iterator/constant_iterator.h>
#include <thrust/iterator/zip_iterator.h>
#include <thrust/functional.h>
#include <thrust/advance.h>
#include <thrust/tuple.h>
#include <thrust/count.h>

/* (1) CUDA-samples helpers */
#define checkCudaErrors(x)                                                                             \
  do {                                                                                                 \
    cudaError_t e_ = (x);                                                                              \
    if (e_ != cudaSuccess) {                                                                           \
      fprintf(stderr, "[ref] CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);   \
      exit(1);                                                                                         \
    }                                                                                                  \
  } while (0)
#define getLastCudaError(msg)                                                                          \
  do {                                                                                                 \
    cudaError_t e_ = cudaGetLastError();                                                               \
    if (e_ != cudaSuccess) {                                                                           \
      fprintf(stderr, "[ref] %s: CUDA error %s at %s:%d\n", msg, cudaGetErrorString(e_), __FILE__,     \
              __LINE__);                                                                               \
      exit(1);                                                                                         \
    }                                                                                                  \
  } while (0)

/* (2) legacy texture references on top of texture objects */
template <class T, int Dim = 1, cudaTextureReadMode Mode = cudaReadModeElementType>
struct ref_texture_shim {
  cudaTextureAddressMode addressMode[3];
  cudaTextureFilterMode filterMode;
  int normalized;
  cudaTextureObject_t obj;
};

template <class T, int D, cudaTextureReadMode M>
static inline cudaError_t cudaBindTextureToArray(ref_texture_shim<T, D, M>& t, cudaArray_t arr,
                                                 const cudaChannelFormatDesc&) {
  if (t.obj) { cudaDestroyTextureObject(t.obj); t.obj = 0; }
  cudaResourceDesc rd = {};
  rd.resType = cudaResourceTypeArray;
  rd.res.array.array = arr;
  cudaTextureDesc td = {};
  for (int i = 0; i < 3; ++i) td.addressMode[i] = t.addressMode[i];
  td.filterMode = t.filterMode;
  td.readMode = M;
  td.normalizedCoords = t.normalized;
  cudaTextureObject_t o = 0;
  cudaError_t e = cudaCreateTextureObject(&o, &rd, &td, nullptr);
  t.obj = o;
  return e;
}
template <class T, int D, cudaTextureReadMode M>
static inline cudaError_t cudaBindTexture(size_t* offset, ref_texture_shim<T, D, M>& t, const void* ptr,
                                          const cudaChannelFormatDesc& desc, size_t bytes) {
  if (offset) *offset = 0;
  if (t.obj) { cudaDestroyTextureObject(t.obj); t.obj = 0; }
  cudaResourceDesc rd = {};
  rd.resType = cudaResourceTypeLinear;
  rd.res.linear.devPtr = const_cast<void*>(ptr);
  rd.res.linear.desc = desc;
  rd.res.linear.sizeInBytes = bytes;
  cudaTextureDesc td = {};
  td.readMode = M; /* tex1Dfetch: integer coordinate, point sampled, like the legacy linear-memory binding */
  cudaTextureObject_t o = 0;
  cudaError_t e = cudaCreateTextureObject(&o, &rd, &td, nullptr);
  t.obj = o;
  return e;
}
template <class T, int D, cudaTextureReadMode M>
static inline cudaError_t cudaUnbindTexture(ref_texture_shim<T, D, M>& t) {
  cudaError_t e = cudaSuccess;
  if (t.obj) { e = cudaDestroyTextureObject(t.obj); t.obj = 0; }
  return e;
}
template <class T, int D, cudaTextureReadMode M>
static __device__ __forceinline__ T tex3D(const ref_texture_shim<T, D, M>& t, float x, float y, float z) {
  return tex3D<T>(t.obj, x, y, z);
}
template <class T, int D, cudaTextureReadMode M>
static __device__ __forceinline__ T tex1Dfetch(const ref_texture_shim<T, D, M>& t, int i) {
  return tex1Dfetch<T>(t.obj, i);
}
/* A legacy texture reference is a file-scope object visible to host and device code: __managed__ gives that. */
#define texture __managed__ ref_texture_shim

/* (4) the sources' std::thread branch (USE_BOOST=0) starts a worker with
 *     std::thread(&GPUWorkerCommunicator::execute, std::ref(shared_ptr)), which only MSVC's lax INVOKE accepted;
 *     this operator lets libstdc++'s INVOKE dereference that wrapper (found by ADL). */
#include <memory>
#include <functional>
template <class T>
inline T& operator*(const std::reference_wrapper<std::shared_ptr<T>>& r) { return *r.get(); }

/* (3) thrust API drift */
namespace thrust {
template <class... T>
__host__ __device__ inline tuple<T...> make_tuple(const T&... t) { return tuple<T...>(t...); }
template <class T, class U, class Op>
inline U reduce(device_ptr<T> a, device_ptr<T> b, U init, Op op) {
  return thrust::reduce(thrust::device, a, b, init, op);
}
template <class Tup, class F, class U, class Op>
inline U transform_reduce(zip_iterator<Tup> a, zip_iterator<Tup> b, F f, U init, Op op) {
  return thrust::transform_reduce(thrust::device, a, b, f, init, op);
}
}  // namespace thrust
