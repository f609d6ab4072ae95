#ifndef MOCK_RINTERNALS_H
#define MOCK_RINTERNALS_H
#include "R.h"
typedef struct SEXPREC* SEXP;
typedef ptrdiff_t R_xlen_t;
typedef unsigned int SEXPTYPE;
#define REALSXP 14
#define INTSXP 13
#define VECSXP 19
#define STRSXP 16
extern SEXP R_NilValue, R_NamesSymbol, R_DimSymbol;
SEXP Rf_getAttrib(SEXP, SEXP);
#define getAttrib Rf_getAttrib
R_xlen_t XLENGTH(SEXP);
int TYPEOF(SEXP);
SEXP STRING_ELT(SEXP, R_xlen_t);
SEXP VECTOR_ELT(SEXP, R_xlen_t);
SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP);
const char* CHAR(SEXP);
double* REAL(SEXP);
int* INTEGER(SEXP);
int Rf_asInteger(SEXP);
double Rf_asReal(SEXP);
int Rf_asLogical(SEXP);
#define asInteger Rf_asInteger
#define asReal Rf_asReal
#define asLogical Rf_asLogical
SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
#define PROTECT(x) Rf_protect(x)
#define UNPROTECT(n) Rf_unprotect(n)
SEXP Rf_allocVector(SEXPTYPE, R_xlen_t);
SEXP Rf_allocMatrix(SEXPTYPE, int, int);
SEXP Rf_alloc3DArray(SEXPTYPE, int, int, int);
SEXP Rf_mkNamed(SEXPTYPE, const char**);
SEXP Rf_ScalarReal(double);
#define allocVector Rf_allocVector
#define allocMatrix Rf_allocMatrix
#define alloc3DArray Rf_alloc3DArray
#define mkNamed Rf_mkNamed
#define ScalarReal Rf_ScalarReal
extern SEXP R_GlobalEnv;
SEXP Rf_lang2(SEXP, SEXP);
SEXP Rf_eval(SEXP, SEXP);
SEXP Rf_coerceVector(SEXP, SEXPTYPE);
#define lang2 Rf_lang2
#define eval Rf_eval
#define coerceVector Rf_coerceVector
#endif
