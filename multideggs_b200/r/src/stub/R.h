/* stub of <R.h> for syntax checking only (see Rinternals.h in this directory) */
#include <stdio.h>
