// Stand-in (oracle/_ref build, test infrastructure): everything lives in the shim's ggml.h.
#pragma once
#include "ggml.h"
