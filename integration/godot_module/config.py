# Godot module "audio_spatializer_gpu": the B200 batched mix path behind the godot-audio-spatializer plugin API.
# Drop this directory into <godot>/modules/audio_spatializer_gpu next to the reference module
# (<godot>/modules/audio_spatializer, BuzzLord/godot-audio-spatializer) and build with scons as usual
# (reference README.md:12; reference config.py:13-27 is the twin of this file).


def can_build(env, platform):
    # needs the reference module (its plugin base classes) and a Linux host with libgas_b200.so
    return platform == "linuxbsd" and not env.get("disable_3d", False)


def configure(env):
    pass


def get_doc_classes():
    return [
        "AudioSpatializer3DGPU",
        "AudioSpatializerInstance3DGPU",
    ]


def get_doc_path():
    return "doc_classes"
