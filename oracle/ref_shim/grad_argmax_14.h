// stands in for the generated Halide header of the same name
#pragma once
#include "ref_pipelines.h"
