/* godot-lite forwarding header (test infrastructure): upstream core/config/project_settings.h */
#pragma once
#include "../../godot_lite_core.h"
