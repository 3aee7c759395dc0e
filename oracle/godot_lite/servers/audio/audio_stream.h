/* godot-lite forwarding header (test infrastructure): upstream servers/audio/audio_stream.h */
#pragma once
#include "../../godot_lite_audio.h"
