// shim: include-path compatibility (the LM layer lives in schur_LM_linear_system.h of this shim)
#pragma once
#include "schur_LM_linear_system.h"
