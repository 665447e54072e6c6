// shim: see ../shim_algo.h (the reference sources include this path)
#pragma once
#include "../shim_algo.h"
