/* minimal stand-in for syntax checking only */
#include <stddef.h>
#include <stdint.h>
void error(const char*, ...);
char* R_alloc(size_t, int);
