/* godot-lite forwarding header (test infrastructure): upstream servers/audio/audio_filter_sw.h */
#pragma once
#include "../../godot_lite_audio.h"
