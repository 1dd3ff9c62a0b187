/* stand-in for <gsl/gsl_matrix.h> (see shim/glib.h) */
#ifndef MMB_SHIM_GSL_MATRIX_H
#define MMB_SHIM_GSL_MATRIX_H
typedef struct gsl_matrix gsl_matrix;
typedef struct gsl_vector gsl_vector;
#endif
