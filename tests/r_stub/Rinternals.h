typedef struct SEXPREC* SEXP; typedef ptrdiff_t R_xlen_t; typedef void* (*DL_FUNC)();
extern SEXP R_NilValue; extern double NA_REAL;
#define REALSXP 14
#define INTSXP 13
#define VECSXP 19
#define TRUE 1
#define FALSE 0
SEXP PROTECT(SEXP); void UNPROTECT(int); SEXP allocVector(int, R_xlen_t); SEXP allocMatrix(int, int, int);
int TYPEOF(SEXP); double* REAL(SEXP); int* INTEGER(SEXP); int LENGTH(SEXP); int nrows(SEXP); int ncols(SEXP); int isNull(SEXP); int asInteger(SEXP); int asLogical(SEXP); double asReal(SEXP);
SEXP ScalarInteger(int); SEXP ScalarReal(double); SEXP VECTOR_ELT(SEXP, R_xlen_t); SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP);
void* R_ExternalPtrAddr(SEXP); void R_ClearExternalPtr(SEXP); SEXP R_MakeExternalPtr(void*, SEXP, SEXP);
typedef void (*R_CFinalizer_t)(SEXP); void R_RegisterCFinalizerEx(SEXP, R_CFinalizer_t, int);
