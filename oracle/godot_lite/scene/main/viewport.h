/* godot-lite forwarding header (test infrastructure): upstream scene/main/viewport.h */
#pragma once
#include "../../godot_lite_scene.h"
