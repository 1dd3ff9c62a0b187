/* Stand-in for <glib.h>: just the names the reference's headers and backends/cuda.c use, so that
 * integration/check.sh can type-check cuda.c against the reference's own headers on a machine
 * without glib (this container).  A real build uses the real glib. */
#ifndef MMB_SHIM_GLIB_H
#define MMB_SHIM_GLIB_H
#include <assert.h>
#include <stdarg.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
typedef int gboolean;
typedef int gint;
typedef unsigned int guint;
typedef char gchar;
typedef unsigned char guchar;
typedef long glong;
typedef unsigned long gulong;
typedef float gfloat;
typedef double gdouble;
typedef void *gpointer;
typedef const void *gconstpointer;
typedef unsigned int guint32;
typedef int gint32;
typedef unsigned long gsize;
typedef unsigned long long guint64;
typedef long long gint64;
typedef struct _GMutex GMutex;
typedef struct _GCond GCond;
typedef struct _GList GList;
typedef struct _GSList GSList;
typedef struct _GHashTable GHashTable;
typedef struct _GString GString;
typedef struct _GArray GArray;
typedef struct _GPtrArray GPtrArray;
typedef struct _GThread GThread;
typedef struct _GThreadPool GThreadPool;
typedef struct _GError GError;
typedef struct _GIOChannel GIOChannel;
typedef unsigned int (*GHashFunc)(gconstpointer);
typedef gboolean (*GEqualFunc)(gconstpointer, gconstpointer);
#ifndef TRUE
#define TRUE 1
#define FALSE 0
#endif
#define G_ASCII_DTOSTR_BUF_SIZE 39
#define g_assert(x) assert(x)
#define g_assert_not_reached() assert(0)
#define g_new0(type, n) ((type *)calloc((n), sizeof(type)))
#define g_new(type, n) ((type *)malloc((n) * sizeof(type)))
#define g_free free
#define g_malloc malloc
#define g_malloc0(n) calloc(1, (n))
#define g_strdup strdup
void g_warning(const gchar *format, ...);
void g_print(const gchar *format, ...);
gchar *g_ascii_formatd(gchar *buffer, gint buf_len, const gchar *format, gdouble d);
gchar *g_ascii_dtostr(gchar *buffer, gint buf_len, gdouble d);
GHashTable *g_hash_table_new(GHashFunc hash_func, GEqualFunc key_equal_func);
void g_hash_table_insert(GHashTable *hash_table, gpointer key, gpointer value);
gpointer g_hash_table_lookup(GHashTable *hash_table, gconstpointer key);
gboolean g_hash_table_remove(GHashTable *hash_table, gconstpointer key);
guint g_hash_table_size(GHashTable *hash_table);
void g_hash_table_destroy(GHashTable *hash_table);
guint g_direct_hash(gconstpointer v);
gboolean g_direct_equal(gconstpointer a, gconstpointer b);
#define GINT_TO_POINTER(i) ((gpointer)(glong)(i))
#define GPOINTER_TO_INT(p) ((gint)(glong)(p))
#endif
