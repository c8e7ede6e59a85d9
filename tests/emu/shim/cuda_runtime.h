// TEST INFRASTRUCTURE (tests/emu): what `#include <cuda_runtime.h>` of the kernel headers resolves to in the
// host-compiled build of the kernels.
#pragma once
#include "../cuda_emu.h"
