/* oracle/shim/ml.h -- see cxcore.h (test infrastructure: lets the reference sources compile unchanged). */
#include "cxcore.h"
