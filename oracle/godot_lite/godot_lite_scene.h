/*
 * godot_lite_scene.h — stand-ins for the scene / physics classes the module queries
 * (scene/3d/{node_3d,camera_3d,audio_listener_3d,velocity_tracker_3d,physics/area_3d}.h,
 * scene/main/viewport.h, servers/physics_server_3d.h).  TEST INFRASTRUCTURE ONLY (see
 * godot_lite_core.h).  These carry no behaviour of their own: the harness stores the query RESULTS
 * (transforms, velocities, the overlapping Area3D, closest points) and the module's code reads them
 * back through the upstream-shaped getters.
 */
#pragma once

#include "godot_lite_core.h"

class Viewport;
class Camera3D;
class AudioListener3D;
class World3D;
class SceneTree;

class RID {
public:
	uint64_t id = 0;
	bool operator==(const RID &o) const { return id == o.id; }
};

class Node : public Object {
	GDCLASS(Node, Object);

public:
	enum {
		NOTIFICATION_ENTER_TREE = 10,
		NOTIFICATION_EXIT_TREE = 11,
		NOTIFICATION_READY = 13,
		NOTIFICATION_PAUSED = 14,
		NOTIFICATION_UNPAUSED = 15,
		NOTIFICATION_PHYSICS_PROCESS = 16,
		NOTIFICATION_PROCESS = 17,
		NOTIFICATION_INTERNAL_PROCESS = 25,
		NOTIFICATION_INTERNAL_PHYSICS_PROCESS = 26,
		NOTIFICATION_PREDELETE = 1,
		NOTIFICATION_SUSPENDED = 9003,
		NOTIFICATION_UNSUSPENDED = 9004,
	};
	/* godot-lite state, set by the harness */
	bool gl_inside_tree = false;
	bool gl_physics_process_internal = false;
	Viewport *gl_viewport = nullptr;
	SceneTree *gl_tree = nullptr;

	bool is_inside_tree() const { return gl_inside_tree; }
	bool can_process() const { return true; }
	void set_physics_process_internal(bool p_enabled) { gl_physics_process_internal = p_enabled; }
	Viewport *get_viewport() const { return gl_viewport; }
	SceneTree *get_tree() const { return gl_tree; }
};

class SceneTree : public Object {
	GDCLASS(SceneTree, Object);

public:
	bool is_paused() const { return false; }
};

class Node3D : public Node {
	GDCLASS(Node3D, Node);

public:
	enum {
		NOTIFICATION_TRANSFORM_CHANGED = 2000,
	};
	Transform3D gl_global_transform;
	Ref<World3D> gl_world_3d;
	bool gl_notify_transform = false;

	Transform3D get_global_transform() const { return gl_global_transform; }
	Ref<World3D> get_world_3d() const;
	void set_notify_transform(bool p_enabled) { gl_notify_transform = p_enabled; }
	void set_disable_scale(bool) {}
};

class World3D : public Resource {
	GDCLASS(World3D, Resource);

public:
	HashSet<Camera3D *> gl_cameras;
	RID gl_space;
	const HashSet<Camera3D *> &get_cameras() const { return gl_cameras; }
	RID get_space() const { return gl_space; }
};
inline Ref<World3D> Node3D::get_world_3d() const { return gl_world_3d; }

class AudioListener3D : public Node3D {
	GDCLASS(AudioListener3D, Node3D);

public:
	Vector3 gl_velocity;
	Vector3 get_doppler_tracked_velocity() const { return gl_velocity; }
};

class Camera3D : public Node3D {
	GDCLASS(Camera3D, Node3D);

public:
	Vector3 gl_velocity;
	Vector3 get_doppler_tracked_velocity() const { return gl_velocity; }
};

class Viewport : public Node {
	GDCLASS(Viewport, Node);

public:
	Camera3D *gl_camera_3d = nullptr;
	AudioListener3D *gl_audio_listener_3d = nullptr;
	bool gl_is_audio_listener_3d = true;
	Camera3D *get_camera_3d() const { return gl_camera_3d; }
	bool is_audio_listener_3d() const { return gl_is_audio_listener_3d; }
	AudioListener3D *get_audio_listener_3d() const { return gl_audio_listener_3d; }
};

class VelocityTracker3D : public RefCounted {
	GDCLASS(VelocityTracker3D, RefCounted);

public:
	Vector3 gl_velocity;
	bool gl_track_physics_step = false;
	void set_track_physics_step(bool p_track_physics_step) { gl_track_physics_step = p_track_physics_step; }
	void reset(const Vector3 &) {}
	void update_position(const Vector3 &) {}
	Vector3 get_tracked_linear_velocity() const { return gl_velocity; }
};

class CollisionObject3D : public Node3D {
	GDCLASS(CollisionObject3D, Node3D);

public:
	RID gl_rid;
	RID get_rid() const { return gl_rid; }
};

class Area3D : public CollisionObject3D {
	GDCLASS(Area3D, CollisionObject3D);

public:
	bool gl_audio_bus_override = false;
	StringName gl_audio_bus_name = StringName("Master");
	bool gl_use_reverb_bus = false;
	StringName gl_reverb_bus_name = StringName("Master");
	float gl_reverb_amount = 0.0;
	float gl_reverb_uniformity = 0.0;
	bool is_overriding_audio_bus() const { return gl_audio_bus_override; }
	StringName get_audio_bus_name() const { return gl_audio_bus_name; }
	bool is_using_reverb_bus() const { return gl_use_reverb_bus; }
	StringName get_reverb_bus_name() const { return gl_reverb_bus_name; }
	float get_reverb_amount() const { return gl_reverb_amount; }
	float get_reverb_uniformity() const { return gl_reverb_uniformity; }
};

class PhysicsDirectSpaceState3D : public Object {
	GDCLASS(PhysicsDirectSpaceState3D, Object);

public:
	struct ShapeResult {
		RID rid;
		uint64_t collider_id = 0;
		Object *collider = nullptr;
		int shape = 0;
	};
	struct PointParameters {
		Vector3 position;
		uint32_t collision_mask = UINT32_MAX;
		bool collide_with_bodies = true;
		bool collide_with_areas = false;
	};
	/* the harness answers the two queries: which areas overlap the point, and the closest point of an
	 * area's volume to a listener */
	std::function<int(const PointParameters &, ShapeResult *, int)> gl_intersect_point;
	std::function<Vector3(RID, const Vector3 &)> gl_closest_point;
	int intersect_point(const PointParameters &p_parameters, ShapeResult *r_results, int p_result_max) {
		return gl_intersect_point ? gl_intersect_point(p_parameters, r_results, p_result_max) : 0;
	}
	Vector3 get_closest_point_to_object_volume(RID p_object, const Vector3 p_point) const {
		return gl_closest_point ? gl_closest_point(p_object, p_point) : Vector3();
	}
};

class PhysicsServer3D : public Object {
	GDCLASS(PhysicsServer3D, Object);

public:
	PhysicsDirectSpaceState3D gl_space_state;
	static PhysicsServer3D *get_singleton() {
		static PhysicsServer3D s;
		return &s;
	}
	PhysicsDirectSpaceState3D *space_get_direct_state(RID) { return &gl_space_state; }
};
