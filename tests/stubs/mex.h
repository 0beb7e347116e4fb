// tests/stubs/mex.h -- minimal stand-in for MATLAB's mex.h: declarations only, used by tests/test_abi.py to syntax-check matlab/emb_mex.cpp
#include <cstddef>
#include <cstdint>
typedef struct mxArray_tag mxArray; typedef size_t mwSize;
enum mxClassID { mxUINT8_CLASS, mxUINT64_CLASS, mxINT8_CLASS, mxUINT16_CLASS, mxINT16_CLASS, mxSINGLE_CLASS, mxDOUBLE_CLASS }; enum mxComplexity { mxREAL };
bool mxIsUint64(const mxArray*); size_t mxGetNumberOfElements(const mxArray*); void* mxGetData(const mxArray*); bool mxIsStruct(const mxArray*);
mxArray* mxGetField(const mxArray*, int, const char*); bool mxIsEmpty(const mxArray*); double mxGetScalar(const mxArray*); double* mxGetPr(const mxArray*);
size_t mxGetM(const mxArray*); bool mxIsChar(const mxArray*); int mxGetString(const mxArray*, char*, size_t);
mxArray* mxCreateNumericMatrix(size_t, size_t, mxClassID, mxComplexity); mxArray* mxCreateStructMatrix(size_t, size_t, int, const char**);
void mxSetField(mxArray*, int, const char*, mxArray*); mxArray* mxCreateDoubleScalar(double); mxArray* mxCreateDoubleMatrix(size_t, size_t, mxComplexity);
mxArray* mxCreateNumericArray(int, const mwSize*, mxClassID, mxComplexity); void mxDestroyArray(mxArray*);
[[noreturn]] void mexErrMsgIdAndTxt(const char*, const char*, ...);
typedef bool mxLogical; bool mxIsLogical(const mxArray*); mxLogical* mxGetLogicals(const mxArray*); mxArray* mxGetCell(const mxArray*, size_t);
mxArray* mxCreateCellMatrix(size_t, size_t); void mxSetCell(mxArray*, size_t, mxArray*);
