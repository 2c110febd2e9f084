// symmetric Gram kernels for d = 1 .. 4 (see pncov_gram_sym.inc)
#define PN_SYM_HALF 0
#include "pncov_gram_sym.inc"
