/* Stand-in for the reference's GENERATED compiler_types.h (ops.lisp:360-400 make-types-file writes it from the *types*
 * list, ops.lisp:37-66); only what headers need to parse: the type numbers in list order and the runtime value union. */
#ifndef MMB_SHIM_COMPILER_TYPES_H
#define MMB_SHIM_COMPILER_TYPES_H
#define TYPE_NIL 0
#define TYPE_INT 1
#define TYPE_FLOAT 2
#define TYPE_COMPLEX 3
#define TYPE_COLOR 4
#define TYPE_CURVE 5
#define TYPE_GRADIENT 6
#define TYPE_IMAGE 7
#define TYPE_TUPLE 8
#define TYPE_TREE_VECTOR 9
#define MAX_TYPE TYPE_TREE_VECTOR
#define RUNTIME_VALUE_DECL int int_value; float float_value; float _Complex complex_value; color_t color_value; curve_t * curve_value; gradient_t * gradient_value; image_t * image_value; float * tuple_value; tree_vector_t * tree_vector_value;
#endif
