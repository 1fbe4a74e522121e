/* Hand-written stand-in for the autoconf-generated config.h of the reference
 * (configure.ac:50-67 defaults on an x86-64 box with AVX-512).  Used only to
 * compile the UNMODIFIED reference sources into oracle/_ref/ (test
 * infrastructure, see oracle/Makefile). */
#pragma once
#define VERSION "1.7"
#define ENABLE_X86_SIMD 1
#define ENABLE_AVX512 1
#define HAVE_FUNC_ATTRIBUTE_IFUNC 1
