/* Minimal stand-in for R's <Rinternals.h>: ONLY for `gcc -fsyntax-only` of rcall.c in an image without R.
 * Declarations follow the public R API (Writing R Extensions); nothing here is linked or shipped to users. */
#ifndef MDEGGS_STUB_RINTERNALS_H
#define MDEGGS_STUB_RINTERNALS_H
#include <stddef.h>
typedef struct SEXPREC* SEXP;
typedef ptrdiff_t R_xlen_t;
typedef unsigned char Rbyte;
typedef enum { FALSE = 0, TRUE } Rboolean;
#define INTSXP 13
#define REALSXP 14
#define STRSXP 16
#define VECSXP 19
#define RAWSXP 24
extern SEXP R_NilValue, R_NamesSymbol;
extern double R_NaReal;
#define NA_REAL R_NaReal
int TYPEOF(SEXP);
double* REAL(SEXP);
int* INTEGER(SEXP);
Rbyte* RAW(SEXP);
R_xlen_t XLENGTH(SEXP);
SEXP Rf_allocVector(unsigned int, R_xlen_t);
SEXP Rf_allocMatrix(unsigned int, int, int);
SEXP Rf_alloc3DArray(unsigned int, int, int, int);
SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)
void Rf_error(const char*, ...);
int Rf_asInteger(SEXP);
int Rf_asLogical(SEXP);
int Rf_nrows(SEXP);
int Rf_ncols(SEXP);
Rboolean Rf_isMatrix(SEXP);
SEXP Rf_setAttrib(SEXP, SEXP, SEXP);
SEXP Rf_mkChar(const char*);
void SET_STRING_ELT(SEXP, R_xlen_t, SEXP);
SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP);
SEXP VECTOR_ELT(SEXP, R_xlen_t);
SEXP Rf_mkString(const char*);
SEXP R_MakeExternalPtr(void*, SEXP, SEXP);
void* R_ExternalPtrAddr(SEXP);
void R_ClearExternalPtr(SEXP);
typedef void (*R_CFinalizer_t)(SEXP);
void R_RegisterCFinalizerEx(SEXP, R_CFinalizer_t, Rboolean);
void R_PreserveObject(SEXP);
void R_ReleaseObject(SEXP);
char* R_alloc(size_t, int);
#endif
