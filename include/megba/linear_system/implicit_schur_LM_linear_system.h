// include-path compatibility with the reference (include/linear_system/implicit_schur_LM_linear_system.h)
#pragma once
#include "schur_LM_linear_system.h"
