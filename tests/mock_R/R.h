/* Minimal stand-in for R's C API headers: DECLARATIONS ONLY, so that r/bssm_b200_shim.c can be
 * syntax- and type-checked with gcc in an image without R (tests/test_abi.py).  Not R's headers,
 * nothing here is ever linked or executed. */
#ifndef MOCK_R_H
#define MOCK_R_H
#include <stddef.h>
void Rf_error(const char*, ...);
void Rf_warning(const char*, ...);
#define error Rf_error
#define warning Rf_warning
char* R_alloc(size_t, int);
int R_finite(double);
#define R_FINITE(x) R_finite(x)
#endif
