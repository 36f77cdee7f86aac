/* oracle R shim: intentionally empty (see R.h) */
#include "R.h"
