/* godot-lite forwarding header (test infrastructure): upstream core/variant/dictionary.h */
#pragma once
#include "../../godot_lite_core.h"
