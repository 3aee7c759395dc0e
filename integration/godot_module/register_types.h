// register_types.h — module entry points of audio_spatializer_gpu (twin of the reference's register_types.h:33-36).
#pragma once

#include "modules/register_module_types.h"

void initialize_audio_spatializer_gpu_module(ModuleInitializationLevel p_level);
void uninitialize_audio_spatializer_gpu_module(ModuleInitializationLevel p_level);
