/* godot-lite forwarding header (test infrastructure): upstream core/io/resource.h */
#pragma once
#include "../../godot_lite_core.h"
