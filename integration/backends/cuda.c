/* backends/cuda.c -- binds libmathmap_b200.so behind the reference's backend boundary.
 *
 * A maintainer adds this file to the reference tree next to backends/cc.c and backends/llvm.cpp, links the plug-in
 * with -lmathmap_b200, and replaces the gen_and_load_c_code call in compile_mathmap (mathmap_common.c:551-556) with
 *
 *     mathmap->initfunc = gen_and_load_cuda_code(mathmap, &mathmap->module_info, filter_codes);
 *
 * and unload_c_code in unload_mathmap (mathmap_common.c:280-291) with unload_cuda_code.  free_invocation
 * (mathmap_common.c:304-319) gains one line, cuda_free_invocation(invocation).  Nothing above the boundary changes:
 * the CLI (mathmap_cmdline.c:798-871) and the GIMP plug-in (mathmap.c:1128-1190, 2160-2225) keep calling
 * invocation_new_frame / call_invocation_parallel_and_join, which reach the three functions of mathfuncs_t below.
 *
 * What it does:
 *   gen_and_load_cuda_code  walks filter_code_t** exactly like backends/cc.c:577-589 and prints the optimised IR as
 *                           "mmir 1" text (grammar: mathmap_b200/csrc/ir/ir_text.cpp), hands it to mmb_load_ir
 *   cuda_initfunc           initfunc_t: one mmb_invocation per mathmap_invocation_t (preview and final render each
 *                           have their own, mathmap.c:1142,2166)
 *   cuda_init_frame         binds closure->v.closure.args[] (the uservals) and calls mmb_init_frame
 *   cuda_calc_lines         forwards the slice, frame and invocation parameters to mmb_calc_lines_slice
 *
 * In this repository the file is type-checked against the reference's own headers by integration/check.sh
 * (gcc -fsyntax-only with stand-ins for glib/gtk/gimp/GSL and the two generated headers); it cannot be linked here
 * because the reference itself cannot be built in this environment.
 */
#include <complex.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <glib.h>

#include "../compiler-internals.h"
#include "../mathmap.h"
#include "../drawable.h"

#include <mathmap_b200.h>

initfunc_t gen_and_load_cuda_code (mathmap_t *mathmap, void **module_info, filter_code_t **filter_codes);
void unload_cuda_code (void *module_info);
void cuda_free_invocation (mathmap_invocation_t *invocation);

/*** IR printer ***/

typedef struct
{
    FILE *out;
    GHashTable *compvar_ids;	/* compvar_t* -> id + 1 */
    int num_compvars;
    compvar_t **compvars;	/* in order of first appearance */
} printer_t;

static int
compvar_id (printer_t *pr, compvar_t *compvar)
{
    int id = GPOINTER_TO_INT(g_hash_table_lookup(pr->compvar_ids, compvar));

    if (id == 0)
    {
	pr->compvars = (compvar_t**)realloc(pr->compvars, sizeof(compvar_t*) * (pr->num_compvars + 1));
	pr->compvars[pr->num_compvars] = compvar;
	id = ++pr->num_compvars;
	g_hash_table_insert(pr->compvar_ids, compvar, GINT_TO_POINTER(id));
    }
    return id - 1;
}

static void
put_float (FILE *out, float f)
{
    char buf[G_ASCII_DTOSTR_BUF_SIZE];

    if (f != f)
	fputs("nan", out);
    else if (f > 3.4028234e38f)
	fputs("inf", out);
    else if (f < -3.4028234e38f)
	fputs("-inf", out);
    else
	fputs(g_ascii_formatd(buf, sizeof(buf), "%.9g", f), out);	/* locale-independent, round-trips a float */
}

static void
put_value (printer_t *pr, value_t *value)
{
    if (value->index < 0)	/* never assigned: reads as zero (backends/cc.c:71) */
	fprintf(pr->out, " %%%d.u", compvar_id(pr, value->compvar));
    else
	fprintf(pr->out, " %%%d.%d", compvar_id(pr, value->compvar), value->index);
}

static void
put_primary (printer_t *pr, primary_t *primary)
{
    if (primary->kind == PRIMARY_VALUE)
    {
	put_value(pr, primary->v.value);
	return;
    }
    g_assert(primary->kind == PRIMARY_CONST);
    switch (primary->const_type)
    {
	case TYPE_INT :
	    fprintf(pr->out, " i:%d", primary->v.constant.int_value);
	    break;
	case TYPE_FLOAT :
	    fputs(" f:", pr->out);
	    put_float(pr->out, primary->v.constant.float_value);
	    break;
	case TYPE_COMPLEX :
	    fputs(" c:", pr->out);
	    put_float(pr->out, crealf(primary->v.constant.complex_value));
	    fputs(",", pr->out);
	    put_float(pr->out, cimagf(primary->v.constant.complex_value));
	    break;
	case TYPE_COLOR :
	    fprintf(pr->out, " k:%u", (unsigned int)primary->v.constant.color_value);
	    break;
	default :
	    g_assert_not_reached();
    }
}

static void
put_primaries (printer_t *pr, primary_t *args, int num)
{
    int i;

    for (i = 0; i < num; ++i)
	put_primary(pr, &args[i]);
}

static void
put_rhs (printer_t *pr, rhs_t *rhs)
{
    switch (rhs->kind)
    {
	case RHS_PRIMARY :
	    put_primary(pr, &rhs->v.primary);
	    return;
	case RHS_INTERNAL :
	    fprintf(pr->out, " (internal %s)", rhs->v.internal->name);
	    return;
	case RHS_OP :
	    fprintf(pr->out, " (op %s", rhs->v.op.op->name);
	    put_primaries(pr, rhs->v.op.args, rhs->v.op.op->num_args);
	    break;
	case RHS_FILTER :
	    /* uservals, x, y, t (compiler_num_filter_args) */
	    fprintf(pr->out, " (filter %s", rhs->v.filter.filter->name);
	    put_primaries(pr, rhs->v.filter.args, compiler_num_filter_args(rhs->v.filter.filter));
	    break;
	case RHS_CLOSURE :
	    /* uservals only (backends/cc.c:160) */
	    fprintf(pr->out, " (closure %s", rhs->v.closure.filter->name);
	    put_primaries(pr, rhs->v.closure.args, compiler_num_filter_args(rhs->v.closure.filter) - 3);
	    break;
	case RHS_TUPLE :
	    fputs(" (tuple", pr->out);
	    put_primaries(pr, rhs->v.tuple.args, rhs->v.tuple.length);
	    break;
	case RHS_TREE_VECTOR :
	    fputs(" (tree-vector", pr->out);
	    put_primaries(pr, rhs->v.tuple.args, rhs->v.tuple.length);
	    break;
	default :
	    g_assert_not_reached();
    }
    fputs(")", pr->out);
}

/* const bits as the reference computed them (value_t.const_type, internals.h:31-34); the level that follows is
   recomputed by mmb_load_ir, 0 is a placeholder */
static void
put_lhs (printer_t *pr, value_t *lhs)
{
    put_value(pr, lhs);
    fprintf(pr->out, " %u 0", (unsigned int)lhs->const_type);
}

static void
put_phis (printer_t *pr, statement_t *phi)
{
    fputs(" (phis", pr->out);
    for (; phi != NULL; phi = phi->next)
    {
	if (phi->kind == STMT_NIL)
	    continue;
	g_assert(phi->kind == STMT_PHI_ASSIGN);
	fputs(" (phi", pr->out);
	put_lhs(pr, phi->v.assign.lhs);
	put_rhs(pr, phi->v.assign.rhs);
	put_rhs(pr, phi->v.assign.rhs2);
	fputs(")", pr->out);
    }
    fputs(")", pr->out);
}

static void
put_stmts (printer_t *pr, statement_t *stmt)
{
    for (; stmt != NULL; stmt = stmt->next)
	switch (stmt->kind)
	{
	    case STMT_NIL :
		break;
	    case STMT_ASSIGN :
		fputs("(assign", pr->out);
		put_lhs(pr, stmt->v.assign.lhs);
		put_rhs(pr, stmt->v.assign.rhs);
		fputs(")\n", pr->out);
		break;
	    case STMT_IF_COND :
		fputs("(if", pr->out);
		put_rhs(pr, stmt->v.if_cond.condition);
		fputs(" 0 (\n", pr->out);
		put_stmts(pr, stmt->v.if_cond.consequent);
		fputs(") (\n", pr->out);
		put_stmts(pr, stmt->v.if_cond.alternative);
		fputs(")", pr->out);
		put_phis(pr, stmt->v.if_cond.exit);
		fputs(")\n", pr->out);
		break;
	    case STMT_WHILE_LOOP :
		fputs("(while", pr->out);
		put_phis(pr, stmt->v.while_loop.entry);
		put_rhs(pr, stmt->v.while_loop.invariant);
		fputs(" 0 (\n", pr->out);
		put_stmts(pr, stmt->v.while_loop.body);
		fputs("))\n", pr->out);
		break;
	    default :
		g_assert_not_reached();
	}
}

static const char*
type_name (type_t type)
{
    static const char *names[] = { "nil", "int", "float", "complex", "color", "curve", "gradient", "image", "tuple", "tree_vector" };

    g_assert(type >= 0 && type <= MAX_TYPE);
    return names[type];
}

static void
put_string (FILE *out, const char *s)
{
    fputc('"', out);
    for (; *s != 0; ++s)
    {
	if (*s == '"' || *s == '\\')
	    fputc('\\', out);
	fputc(*s, out);
    }
    fputc('"', out);
}

static void
put_uservals (FILE *out, userval_info_t *info)
{
    fputs(" (uservals", out);
    for (; info != NULL; info = info->next)
    {
	switch (info->type)
	{
	    case USERVAL_INT_CONST :
		fputs(" (int ", out);
		put_string(out, info->name);
		fprintf(out, " %d %d %d", info->v.int_const.min, info->v.int_const.max, info->v.int_const.default_value);
		break;
	    case USERVAL_FLOAT_CONST :
		fputs(" (float ", out);
		put_string(out, info->name);
		fputs(" ", out); put_float(out, info->v.float_const.min);
		fputs(" ", out); put_float(out, info->v.float_const.max);
		fputs(" ", out); put_float(out, info->v.float_const.default_value);
		break;
	    case USERVAL_BOOL_CONST :
		fputs(" (bool ", out);
		put_string(out, info->name);
		fprintf(out, " %d", info->v.bool_const.default_value);
		break;
	    case USERVAL_COLOR :
		fputs(" (color ", out);
		put_string(out, info->name);
		break;
	    case USERVAL_CURVE :
		fputs(" (curve ", out);
		put_string(out, info->name);
		break;
	    case USERVAL_GRADIENT :
		fputs(" (gradient ", out);
		put_string(out, info->name);
		break;
	    case USERVAL_IMAGE :
		fputs(" (image ", out);
		put_string(out, info->name);
		fprintf(out, " %u", info->v.image.flags);
		break;
	    default :
		g_assert_not_reached();
	}
	fputs(")", out);
    }
    fputs(")\n", out);
}

/* (filter NAME (flags ...) (uservals ...) (vars (ID TYPE) ...) (code ...)).  The code is printed first into a
   buffer of its own because the (vars ...) list -- every compvar the code mentions -- precedes it in the text. */
static void
put_filter (FILE *out, filter_code_t *code)
{
    filter_t *filter = code->filter;
    unsigned int flags = filter_flags(filter);
    printer_t pr;
    char *code_text = NULL;
    size_t code_len = 0;
    int i;

    pr.out = open_memstream(&code_text, &code_len);
    pr.compvar_ids = g_hash_table_new(g_direct_hash, g_direct_equal);
    pr.num_compvars = 0;
    pr.compvars = NULL;
    put_stmts(&pr, code->first_stmt);
    fclose(pr.out);

    fprintf(out, "(filter %s (flags%s%s)\n", filter->name,
	    (flags & IMAGE_FLAG_UNIT) ? " unit" : "", (flags & IMAGE_FLAG_SQUARE) ? " square" : "");
    put_uservals(out, filter->userval_infos);
    fputs(" (vars", out);
    for (i = 0; i < pr.num_compvars; ++i)
	fprintf(out, " (%d %s)", i, type_name(pr.compvars[i]->type));	/* tuple lengths are inferred by mmb_load_ir */
    fputs(")\n (code\n", out);
    fwrite(code_text, 1, code_len, out);
    fputs("))\n", out);

    free(code_text);
    free(pr.compvars);
    g_hash_table_destroy(pr.compvar_ids);
}

/*** invocations ***/

/* One mmb_invocation per mathmap_invocation_t.  calc_lines is called concurrently from the band threads of
   call_invocation_parallel (mathmap_common.c:973-1006): launches are serialised per process by one lock, which a
   GPU backend wants anyway (each band is one kernel launch plus a copy). */
typedef struct _binding_t
{
    mathmap_invocation_t *invocation;
    mmb_invocation *inv;
    input_drawable_t **bound_drawables;	/* per userval: the drawable whose pixels are on the device */
    struct _binding_t *next;
} binding_t;

static binding_t *bindings = NULL;
static pthread_mutex_t bindings_mutex = PTHREAD_MUTEX_INITIALIZER;

static binding_t*
lookup_binding (mathmap_invocation_t *invocation)
{
    binding_t *b;

    for (b = bindings; b != NULL; b = b->next)
	if (b->invocation == invocation)
	    return b;
    return NULL;
}

static void
report (const char *what)
{
    g_warning("mathmap_b200: %s: %s", what, mmb_last_error());
}

/* invocation fields that may change between calls (mathmap.c:1168-1169 sets row_stride and output_bpp per region) */
static void
sync_settings (binding_t *b)
{
    mathmap_invocation_t *invocation = b->invocation;

    mmb_set_antialiasing(b->inv, invocation->antialiasing);
    mmb_set_supersampling(b->inv, invocation->supersampling);	/* only the sampler's +0.5; the caller combines slices */
    mmb_set_edge_behaviour(b->inv, invocation->edge_behaviour_x, invocation->edge_behaviour_y,
			   invocation->edge_color_x, invocation->edge_color_y);
    mmb_set_output_bpp(b->inv, invocation->output_bpp);
    mmb_set_render_size(b->inv, invocation->render_width, invocation->render_height);
}

/* The pixels of an input drawable as RGBA8 through the reference's own accessor, which hides GIMP tiles, the
   command line's image cache and movie frames (mathmap.c:1195-1260, mathmap_cmdline.c:131-184). */
static int
upload_drawable (binding_t *b, int index, image_t *image, int frame)
{
    input_drawable_t *drawable = image->v.drawable;
    int width = image->pixel_width, height = image->pixel_height;
    unsigned char *pixels, *p;
    int x, y, result;

    if (b->bound_drawables[index] == drawable)
	return 0;
    pixels = (unsigned char*)malloc((size_t)width * height * 4);
    if (pixels == NULL)
	return -1;
    p = pixels;
    for (y = 0; y < height; ++y)
	for (x = 0; x < width; ++x)
	{
	    color_t c = mathmap_get_pixel(b->invocation, drawable, frame, x, y);

	    p[0] = RED(c); p[1] = GREEN(c); p[2] = BLUE(c); p[3] = ALPHA(c);
	    p += 4;
	}
    result = mmb_set_userval_image_host(b->inv, index, pixels, width, height);
    free(pixels);
    if (result == 0)
	b->bound_drawables[index] = drawable;
    return result;
}

static void
cuda_init_frame (mathmap_frame_t *mmframe, image_t *closure)
{
    mathmap_invocation_t *invocation = mmframe->invocation;
    userval_t *args = closure->v.closure.args;
    userval_info_t *info;
    binding_t *b;
    int ok = 1;

    pthread_mutex_lock(&bindings_mutex);
    b = lookup_binding(invocation);
    g_assert(b != NULL);
    sync_settings(b);
    for (info = invocation->mathmap->main_filter->userval_infos; info != NULL && ok; info = info->next)
    {
	userval_t *arg = &args[info->index];

	switch (info->type)
	{
	    case USERVAL_INT_CONST :
		ok = mmb_set_userval_int(b->inv, info->index, arg->v.int_const) == 0;
		break;
	    case USERVAL_FLOAT_CONST :
		ok = mmb_set_userval_float(b->inv, info->index, arg->v.float_const) == 0;
		break;
	    case USERVAL_BOOL_CONST :
		ok = mmb_set_userval_bool(b->inv, info->index, arg->v.bool_const) == 0;
		break;
	    case USERVAL_COLOR :
		ok = mmb_set_userval_color_packed(b->inv, info->index, arg->v.color.value) == 0;
		break;
	    case USERVAL_CURVE :
		ok = mmb_set_userval_curve(b->inv, info->index, arg->v.curve->values) == 0;
		break;
	    case USERVAL_GRADIENT :
		ok = mmb_set_userval_gradient(b->inv, info->index, arg->v.gradient->values) == 0;
		break;
	    case USERVAL_IMAGE :
		if (arg->v.image != NULL && arg->v.image->type == IMAGE_DRAWABLE)
		    ok = upload_drawable(b, info->index, arg->v.image, mmframe->current_frame) == 0;
		break;
	    default :
		g_assert_not_reached();
	}
    }
    if (ok)
	ok = mmb_init_frame(b->inv, mmframe->current_frame, mmframe->current_t) == 0;
    if (!ok)
	report("init_frame");
    pthread_mutex_unlock(&bindings_mutex);
    mmframe->xy_vars = NULL;
}

static void
cuda_init_slice (mathmap_slice_t *slice, image_t *closure)
{
    /* per-column values (the reference's y_vars, new_template.c.in:339-373) live on the device */
    slice->y_vars = NULL;
}

static void
cuda_calc_lines (mathmap_slice_t *slice, image_t *closure, int first_row, int last_row, void *q, int floatmap)
{
    mathmap_frame_t *mmframe = slice->frame;
    mathmap_invocation_t *invocation = mmframe->invocation;
    mmb_slice s;
    binding_t *b;
    int row;

    s.frame_render_width = mmframe->frame_render_width;
    s.frame_render_height = mmframe->frame_render_height;
    s.region_x = slice->region_x;
    s.region_y = slice->region_y;
    s.region_width = slice->region_width;
    s.region_height = slice->region_height;
    s.sampling_offset_x = slice->sampling_offset_x;
    s.sampling_offset_y = slice->sampling_offset_y;
    s.row_stride = invocation->row_stride;

    pthread_mutex_lock(&bindings_mutex);
    b = lookup_binding(invocation);
    g_assert(b != NULL);
    mmb_set_output_bpp(b->inv, invocation->output_bpp);
    if (mmb_calc_lines_slice(b->inv, &s, first_row, last_row, q, floatmap) != 0)
	report("calc_lines");
    pthread_mutex_unlock(&bindings_mutex);

    /* progress for the plug-in's preview, new_template.c.in:304-305 */
    if (!invocation->supersampling)
    {
	if (first_row < 0)
	    first_row = 0;
	if (last_row > slice->region_y + slice->region_height)
	    last_row = slice->region_y + slice->region_height;
	for (row = first_row - slice->region_y; row < last_row - slice->region_y; ++row)
	    invocation->rows_finished[row] = 1;
    }
}

static mathfuncs_t
cuda_initfunc (mathmap_invocation_t *invocation)
{
    mathfuncs_t funcs;
    binding_t *b = g_new0(binding_t, 1);
    int num_uservals = invocation->mathmap->main_filter->num_uservals;

    memset(&funcs, 0, sizeof(funcs));
    funcs.init_frame = cuda_init_frame;
    funcs.init_slice = cuda_init_slice;
    funcs.calc_lines = cuda_calc_lines;

    b->invocation = invocation;
    b->inv = mmb_invoke((mmb_module*)invocation->mathmap->module_info, invocation->img_width, invocation->img_height, 0);
    if (b->inv == NULL)
	report("invoke");
    b->bound_drawables = g_new0(input_drawable_t*, num_uservals > 0 ? num_uservals : 1);

    pthread_mutex_lock(&bindings_mutex);
    b->next = bindings;
    bindings = b;
    pthread_mutex_unlock(&bindings_mutex);

    return funcs;
}

/* called from free_invocation (mathmap_common.c:304) */
void
cuda_free_invocation (mathmap_invocation_t *invocation)
{
    binding_t **bp, *b;

    pthread_mutex_lock(&bindings_mutex);
    for (bp = &bindings; *bp != NULL; bp = &(*bp)->next)
	if ((*bp)->invocation == invocation)
	{
	    b = *bp;
	    *bp = b->next;
	    mmb_invocation_free(b->inv);
	    g_free(b->bound_drawables);
	    g_free(b);
	    break;
	}
    pthread_mutex_unlock(&bindings_mutex);
}

/*** compiling and loading/unloading ***/

initfunc_t
gen_and_load_cuda_code (mathmap_t *mathmap, void **module_info, filter_code_t **filter_codes)
{
    char *text = NULL;
    size_t len = 0;
    FILE *out = open_memstream(&text, &len);
    filter_t *filter;
    mmb_module *module;
    int i;

    if (out == NULL)
    {
	sprintf(error_string, "Could not allocate the IR text.");
	return 0;
    }
    fputs("(mmir 1\n", out);
    /* filter_codes is index-aligned with mathmap->filters; native filters have no code (backends/cc.c:577-589) */
    for (i = 0, filter = mathmap->filters; filter != NULL; ++i, filter = filter->next)
    {
	if (filter->kind != FILTER_MATHMAP)
	    continue;
	g_assert(filter_codes[i]->filter == filter);
	put_filter(out, filter_codes[i]);
    }
    fprintf(out, "(main %s))\n", mathmap->main_filter->name);
    fclose(out);

    /* the IR lives in compiler pools that are freed right after we return (mathmap_common.c:558): the module keeps
       no pointer into it, only what it parsed from the text */
    module = mmb_load_ir(text);
    free(text);
    if (module == NULL)
    {
	snprintf(error_string, 1024, "%s", mmb_last_error());
	return 0;
    }
    *module_info = module;
    return cuda_initfunc;
}

void
unload_cuda_code (void *module_info)
{
    mmb_module_free((mmb_module*)module_info);
}
