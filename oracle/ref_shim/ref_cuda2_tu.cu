/* TEST INFRASTRUCTURE (oracle/): translation unit that compiles the UNMODIFIED reference file
 * reconstruction_cuda2.cu (path given by -DREF_CUDA2_CU=...).  The reference says `using namespace thrust;` at
 * file scope (cuda2.cu:38); with today's CCCL that makes the identifier `cuda` ambiguous (`::cuda` vs
 * `thrust::cuda`) inside nvcc's generated kernel stubs.  We therefore hand the file a curated namespace that
 * re-exports exactly the thrust names it uses and nothing else.  No reference code is copied. */
#include "helper_cuda.h"
#include <thrust/system_error.h>
#include <cufft.h>

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
#define thrust ref_thrust
#define REF_STR2(x) #x
#define REF_STR(x) REF_STR2(x)
#include REF_STR(REF_CUDA2_CU)
#undef thrust
#include "ref_capi.cu"
