/* godot-lite forwarding header (test infrastructure): upstream servers/audio/audio_effect.h */
#pragma once
#include "../../godot_lite_audio.h"
