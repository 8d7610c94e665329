/* Minimal stand-in for MATLAB's mex.h: just the declarations matlab/tcmcmc_mex.cu uses, so that the gateway can be
 * syntax-checked where MATLAB is not installed (tests/test_host_logic.py).  Test infrastructure, not a MEX runtime. */
#pragma once
#include <cstddef>
typedef struct mxArray_tag mxArray;
typedef size_t mwSize;
typedef enum { mxREAL, mxCOMPLEX } mxComplexity;
typedef enum { mxDOUBLE_CLASS = 6, mxINT32_CLASS = 12, mxINT64_CLASS = 14 } mxClassID;
extern "C" {
double *mxGetPr(const mxArray *);
double mxGetScalar(const mxArray *);
void *mxGetData(const mxArray *);
size_t mxGetM(const mxArray *);
size_t mxGetN(const mxArray *);
size_t mxGetNumberOfElements(const mxArray *);
mxArray *mxGetField(const mxArray *, size_t, const char *);
void mxSetField(mxArray *, size_t, const char *, mxArray *);
mxArray *mxCreateDoubleMatrix(size_t, size_t, mxComplexity);
mxArray *mxCreateDoubleScalar(double);
mxArray *mxCreateNumericMatrix(size_t, size_t, mxClassID, mxComplexity);
mxArray *mxCreateNumericArray(size_t, const size_t *, mxClassID, mxComplexity);
mxArray *mxCreateStructMatrix(size_t, size_t, int, const char **);
bool mxIsStruct(const mxArray *);
bool mxIsDouble(const mxArray *);
bool mxIsChar(const mxArray *);
bool mxIsEmpty(const mxArray *);
char *mxArrayToString(const mxArray *);
int mxGetString(const mxArray *, char *, size_t);
void mxFree(void *);
void mexErrMsgIdAndTxt(const char *, const char *, ...);
void mexPrintf(const char *, ...);
}
