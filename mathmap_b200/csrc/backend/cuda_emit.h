// IR -> CUDA C.  The counterpart of the reference's IR -> C printer
// (backends/cc.c:73-448) and per-filter skeleton (new_template.c.in), designed
// for the GPU: plain statements instead of GNU statement expressions, tuples in
// registers, no per-pixel heap, frame-constant values passed as a by-value
// uniforms struct computed on the host, one pixel per thread.
#pragma once
#include <map>
#include <string>
#include <vector>

#include "../cabi_module.h"

namespace mmbackend {

struct UniformField {
    const mm::Value *value;
    mm::Type type;
    int tuple_len;
    size_t offset;
    size_t size;
    int image_slot = -1;  // T_IMAGE: index into mm_params::images, fixed at compile time
};

struct FilterKernel {
    const mm::Filter *filter = nullptr;
    std::string kernel_name;           // extern "C" __global__ entry
    std::vector<UniformField> uniforms;
    size_t uniforms_size = 0;          // sizeof(mm_uniforms_<f>) (>= 4)
    std::string row_kernel_name;       // row pre-kernel (empty: none); fills `row_slots` 4-byte arrays of num_rows entries
    int row_slots = 0;
    int filter_index = -1;             // position in the module's filter list: mm_image::closure_filter of its closures
    bool closure_fn = false;           // a device function mm_closure_<f> exists (the filter occurs as a closure value)
    int auto_rows = 1;                 // the most 32x8 tiles a block renders in sequence (1 for kernels with per-pixel loops)
    bool quad = false;                 // a thread renders 4 adjacent pixels of a row: tiles are 128 x 8 pixels (mm_runtime.cuh: quad kernels)
    // The filter's pixel is `img(xy)`: ORIG_VAL of a frame-constant image at the pixel's own coordinates, output as it
    // is (Blur/Gaussian Blur: `blurred(xy)`).  The invocation may then let the producer of that image write the
    // pixels (invocation.cpp: pass-through).  The uniform whose image is sampled, or null.
    const mm::Value *passthrough_image = nullptr;
    // Frame-constant values the pixel kernel branches on (`if (userval)`): condition i is MM_COND_<filter>_<i>(U.v...) in
    // the generated code, which -DMM_SPEC_<filter>_<i>=0|1 turns into a literal, so that NVRTC drops the branch, the
    // untaken side and the constant load (nvrtc_module.cpp: KernelConfig::spec, invocation.cpp: mmb_init_frame).
    std::vector<const mm::Value *> spec_conds;
    std::string spec_prefix;           // "MM_SPEC_<filter>_"
};

struct CudaModuleSource {
    std::string text;                                  // generated part (prepended with the runtime by the NVRTC driver)
    std::map<const mm::Filter *, FilterKernel> kernels;  // every mathmap filter that can be rendered
    bool has_calls = false;  // the module calls filters as device functions (recursion): device stack limit, depth guard (mm_call_overflow)
};

// Throws mm::CompileError when the IR uses something the device cannot do.
CudaModuleSource emit_cuda_module(const mmb_module &m);

}  // namespace mmbackend
