/* stand-in for <gsl/gsl_version.h> (see shim/glib.h) */
#define GSL_MAJOR_VERSION 2
#define GSL_MINOR_VERSION 7
