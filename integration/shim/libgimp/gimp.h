/* stand-in for <libgimp/gimp.h> (see shim/glib.h) */
#ifndef MMB_SHIM_GIMP_H
#define MMB_SHIM_GIMP_H
#include <glib.h>
typedef struct { gdouble r, g, b, a; } GimpRGB;
typedef struct _GimpDrawable GimpDrawable;
typedef struct _GimpTile GimpTile;
typedef struct _GimpPixelRgn GimpPixelRgn;
typedef struct _GimpParam GimpParam;
#endif
