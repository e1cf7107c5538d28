// Minimal stand-in for tensorflow/core/util/gpu_kernel_helper.h (see framework/op_kernel.h): nothing of it is used
// beyond what op_kernel.h's stand-in already declares (the GPU device handle).
#pragma once
#include "tensorflow/core/framework/op_kernel.h"
