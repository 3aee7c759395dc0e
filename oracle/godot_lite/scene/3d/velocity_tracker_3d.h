/* godot-lite forwarding header (test infrastructure): upstream scene/3d/velocity_tracker_3d.h */
#pragma once
#include "../../godot_lite_scene.h"
