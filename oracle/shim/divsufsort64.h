/* Minimal declaration of the libdivsufsort64 interface the reference uses
 * (src/esa.cxx:74, src/esa.h:16,33; src/sequence.cxx:16).  The library itself
 * is absent from this image; oracle/sa_standin.cxx provides the symbol. */
#pragma once
#include <stdint.h>
typedef int64_t saidx64_t;
#ifdef __cplusplus
extern "C" {
#endif
int divsufsort64(const unsigned char *T, saidx64_t *SA, saidx64_t n);
#ifdef __cplusplus
}
#endif
