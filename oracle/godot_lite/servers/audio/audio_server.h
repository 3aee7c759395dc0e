/* godot-lite forwarding header (test infrastructure): upstream servers/audio/audio_server.h */
#pragma once
#include "../../godot_lite_audio.h"
