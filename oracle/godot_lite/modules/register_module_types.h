/* godot-lite forwarding header (test infrastructure): upstream modules/register_module_types.h */
#pragma once
#include "../godot_lite_core.h"
