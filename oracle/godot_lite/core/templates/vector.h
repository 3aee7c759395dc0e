/* godot-lite forwarding header (test infrastructure): upstream core/templates/vector.h */
#pragma once
#include "../../godot_lite_core.h"
