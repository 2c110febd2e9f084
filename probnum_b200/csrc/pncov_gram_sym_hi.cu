// symmetric Gram kernels for d = 5 .. 8 (see pncov_gram_sym.inc)
#define PN_SYM_HALF 1
#include "pncov_gram_sym.inc"
