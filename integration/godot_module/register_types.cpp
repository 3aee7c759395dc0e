// register_types.cpp — registers the GPU-backed spatializer classes (twin of reference register_types.cpp:40-66) and
// owns the lifetime of the device context.
#include "register_types.h"

#include "audio_spatializer_3d_gpu.h"
#include "gas_backend.h"

#include "core/object/class_db.h"

void initialize_audio_spatializer_gpu_module(ModuleInitializationLevel p_level) {
	if (p_level != MODULE_INITIALIZATION_LEVEL_SCENE) {
		return;
	}
	GDREGISTER_CLASS(AudioSpatializer3DGPU);
	GDREGISTER_CLASS(AudioSpatializerInstance3DGPU);
	// the context itself is created lazily by the first instance (AudioServer must be up: speaker mode, mix rate, buses)
}

void uninitialize_audio_spatializer_gpu_module(ModuleInitializationLevel p_level) {
	if (p_level != MODULE_INITIALIZATION_LEVEL_SCENE) {
		return;
	}
	GasBackend::shutdown();
}
