// stubs.cu — intentionally empty: every entry point declared in include/dbx.h is implemented
// (there is no CPU fallback anywhere in the library).
#include "runtime.h"
