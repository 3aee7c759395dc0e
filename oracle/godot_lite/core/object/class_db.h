/* godot-lite forwarding header (test infrastructure): upstream core/object/class_db.h */
#pragma once
#include "../../godot_lite_core.h"
