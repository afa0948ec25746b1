// compat/Halide.h — stands in for <Halide.h> on machines without Halide.  The operator API
// only needs the runtime Buffer type; the generator DSL is not part of this project.
#pragma once
#include "HalideBuffer.h"
