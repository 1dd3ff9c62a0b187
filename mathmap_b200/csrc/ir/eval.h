#pragma once
#include "ir.h"

namespace mm {

// Evaluates a pure scalar op on constant arguments with the reference's host
// semantics.  Returns false when the op cannot be evaluated on the host (images,
// tuples, special functions that live only in the device runtime).
bool eval_op(const OpInfo *op, const Const *args, Const *out);
bool const_is_true(const Const &c);
float const_as_float(const Const &c);
int const_as_int(const Const &c);
std::complex<float> const_as_complex(const Const &c);

}  // namespace mm
