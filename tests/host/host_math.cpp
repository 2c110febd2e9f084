// Host build of pncov_math.cuh for CPU-side validation of the device algorithms
// (tests/test_device_math.py).  Not part of the product.
#define PNCOV_HOST_MATH 1
#include "pncov_math.cuh"

static const double kExp2Tab[PN_EXP2_N] = PNCOV_EXP2_TABLE_INIT;
static const double kLog2Tab[2 * PN_LOG2_N] = PNCOV_LOG2_TABLE_INIT;

extern "C" {
void hm_exp2(const double* u, double* out, long n) { for (long i = 0; i < n; ++i) out[i] = pn_exp2(u[i], kExp2Tab); }
void hm_sqrt(const double* u, double* out, long n) { for (long i = 0; i < n; ++i) out[i] = pn_sqrt(u[i]); }
void hm_rcp(const double* u, double* out, long n) { for (long i = 0; i < n; ++i) out[i] = pn_rcp(u[i]); }
void hm_log2(const double* u, double* out, long n) { for (long i = 0; i < n; ++i) out[i] = pn_log2_ge1(u[i], kLog2Tab); }
}
