/* TEST INFRASTRUCTURE (oracle/): ONE translation unit that compiles the UNMODIFIED reference PVR CUDA sources (volume.cu,
 * reconVolume.cu, patchBased{PSFReconstruction,SimulatePatches,Superresolution,RobustStatistics}_gpu.cu,
 * initPatchBasedRecon_gpu.cu under $(REF_SRC)) plus the harness (ref_pvr_capi.cu) into oracle/_ref/libref_pvr.so.
 * Why one TU, built with -rdc=true: the sources share `extern __constant__ _PSF` (needs relocatable device code), and
 * ReconVolume<T>::getReconValueFromTexture (reconVolume.cu:169-187) returns a reference to a local, which only yields the
 * texture value when it is inlined into the calling kernel as in the reference's own whole-program build (as separate
 * objects it returned a stale stack slot, with device LTO NaN).  `thrust` is narrowed to the names the sources use
 * (`using namespace thrust` vs CCCL's ::cuda, see ref_cuda2_tu.cu); `private` / `protected` are opened so the harness
 * can fill the device structures without the IRTK-based host initialisers.  No reference code is copied. */
#include "helper_cuda.h"
#include <thrust/system_error.h>
#include <fstream>
#include <iostream>

namespace ref_thrust {
using ::thrust::device_ptr;
using ::thrust::device_vector;
using ::thrust::host_vector;
using ::thrust::tuple;
using ::thrust::make_tuple;
using ::thrust::get;
using ::thrust::zip_iterator;
using ::thrust::make_zip_iterator;
using ::thrust::constant_iterator;
using ::thrust::make_constant_iterator;
using ::thrust::copy;
using ::thrust::count;
using ::thrust::count_if;
using ::thrust::fill;
using ::thrust::transform;
using ::thrust::reduce;
using ::thrust::transform_reduce;
using ::thrust::inner_product;
using ::thrust::raw_pointer_cast;
using ::thrust::device_pointer_cast;
using ::thrust::system_error;
using ::thrust::plus;
using ::thrust::minus;
using ::thrust::multiplies;
using ::thrust::divides;
using ::thrust::maximum;
using ::thrust::minimum;
using ::thrust::unary_function;
using ::thrust::binary_function;
}  // namespace ref_thrust
#include "irtk_stub.h"
#define REF_STR2(x) #x
#define REF_STR(x) REF_STR2(x)
#define thrust ref_thrust
#define private public
#define protected public
#include REF_STR(REF_SRC/volume.cu)
#include REF_STR(REF_SRC/reconVolume.cu)
#include REF_STR(REF_SRC/patchBasedPSFReconstruction_gpu.cu)
#include REF_STR(REF_SRC/patchBasedSimulatePatches_gpu.cu)
#include REF_STR(REF_SRC/patchBasedSuperresolution_gpu.cu)
#include REF_STR(REF_SRC/patchBasedRobustStatistics_gpu.cu)
#include REF_STR(REF_SRC/initPatchBasedRecon_gpu.cu)   /* defines _PSF: after the `extern` declarations of the others */
#undef private
#undef protected
#undef thrust
#include "ref_pvr_capi.cu"
