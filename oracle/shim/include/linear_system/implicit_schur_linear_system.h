#pragma once
#include "linear_system/implicit_schur_LM_linear_system.h"
