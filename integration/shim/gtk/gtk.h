/* stand-in for <gtk/gtk.h> (see shim/glib.h): opaque widget types named by the reference's headers */
#ifndef MMB_SHIM_GTK_H
#define MMB_SHIM_GTK_H
#include <glib.h>
typedef struct _GtkWidget GtkWidget;
typedef struct _GtkObject GtkObject;
typedef struct _GtkAdjustment GtkAdjustment;
typedef struct _GtkTreeStore GtkTreeStore;
typedef struct _GtkTreeIter GtkTreeIter;
typedef struct _GtkTextBuffer GtkTextBuffer;
typedef struct _GdkPixbuf GdkPixbuf;
#endif
