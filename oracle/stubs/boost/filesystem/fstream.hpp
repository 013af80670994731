// Last header included by the reference's main.h (main.h:35). TEST INFRASTRUCTURE ONLY.
// The reference seeds cuRAND with clock64() (APD.cu:1270), which makes it irreproducible;
// pinning the seed here (after curand_kernel.h has already been included by main.h:16)
// leaves the reference sources untouched while making the oracle repeatable.
#include <boost/filesystem.hpp>
#ifdef __CUDACC__
static __device__ unsigned long long dvp_oracle_seed = 0x5EEDULL;  // set by the harness per upload
#define clock64() (dvp_oracle_seed)
#endif
