// see op_kernel.h (stand-in headers, test infrastructure)
#pragma once
#include "tensorflow/core/framework/op_kernel.h"
