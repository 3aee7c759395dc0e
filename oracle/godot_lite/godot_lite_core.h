/*
 * godot_lite_core.h — "godot-lite": the smallest stand-in for the Godot engine headers that lets the
 * UNMODIFIED sources of BuzzLord/godot-audio-spatializer (/root/reference/*.cpp) compile and run outside
 * the engine.  TEST INFRASTRUCTURE ONLY (see oracle/gas_oracle.h): it exists so that oracle/_ref can
 * execute the reference's own code and pin the restated oracle against it.
 *
 * What is in here is NOT reference code: it is a from-memory stand-in for upstream Godot 4.x
 * (core/math, core/templates, core/object, core/variant).  Only what the module touches exists.
 * Arithmetic that reaches the audio path (Vector2/3, Basis, Transform3D, Math::*) follows upstream's
 * operation order as recalled; containers only have to behave (value
 * semantics, insertion-ordered HashMap/HashSet/Dictionary, head-inserting SafeList).
 */
#pragma once

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <list>
#include <memory>
#include <mutex>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

/* ---- typedefs / macros (core/typedefs.h, core/math/math_defs.h) --------------------------------- */
typedef float real_t; /* default (single-precision) build */
#define CMP_EPSILON 0.00001
#define _ALWAYS_INLINE_ inline
#define _FORCE_INLINE_ inline
#define Math_PI 3.1415926535897932384626433833
#define Math_TAU 6.2831853071795864769252867666

template <typename T, typename T2>
constexpr auto MAX(const T m_a, const T2 m_b) {
	return m_a > m_b ? m_a : m_b;
}
template <typename T, typename T2>
constexpr auto MIN(const T m_a, const T2 m_b) {
	return m_a < m_b ? m_a : m_b;
}
template <typename T, typename T2, typename T3>
constexpr auto CLAMP(const T m_a, const T2 m_min, const T3 m_max) {
	return m_a < m_min ? m_min : (m_a > m_max ? m_max : m_a);
}

/* ---- error macros (core/error/error_macros.h): log + early return, never throw ------------------- */
namespace godot_lite {
struct ErrorLog {
	int count = 0;
	std::string last;
	bool verbose = false;
};
inline ErrorLog &error_log() {
	static ErrorLog l;
	return l;
}
inline void report_error(const char *what, const char *msg, const char *file, int line) {
	ErrorLog &l = error_log();
	l.count++;
	l.last = std::string(what) + (msg && *msg ? std::string(": ") + msg : std::string()) + " @" + file + ":" + std::to_string(line);
	if (l.verbose) {
		fprintf(stderr, "ERROR: %s\n", l.last.c_str());
	}
}
} // namespace godot_lite

#define GL_MSG_STR(m) ::godot_lite::to_cstr(m)
#define ERR_FAIL_COND(m_cond)                                                     \
	if (m_cond) {                                                                 \
		::godot_lite::report_error("Condition \"" #m_cond "\" is true", "", __FILE__, __LINE__); \
		return;                                                                   \
	} else                                                                        \
		((void)0)
#define ERR_FAIL_COND_MSG(m_cond, m_msg)                                          \
	if (m_cond) {                                                                 \
		::godot_lite::report_error("Condition \"" #m_cond "\" is true", GL_MSG_STR(m_msg), __FILE__, __LINE__); \
		return;                                                                   \
	} else                                                                        \
		((void)0)
#define ERR_FAIL_COND_V(m_cond, m_retval)                                         \
	if (m_cond) {                                                                 \
		::godot_lite::report_error("Condition \"" #m_cond "\" is true", "", __FILE__, __LINE__); \
		return m_retval;                                                          \
	} else                                                                        \
		((void)0)
#define ERR_FAIL_COND_V_MSG(m_cond, m_retval, m_msg)                              \
	if (m_cond) {                                                                 \
		::godot_lite::report_error("Condition \"" #m_cond "\" is true", GL_MSG_STR(m_msg), __FILE__, __LINE__); \
		return m_retval;                                                          \
	} else                                                                        \
		((void)0)
#define ERR_FAIL_INDEX(m_index, m_size)                                           \
	if ((m_index) < 0 || (m_index) >= (m_size)) {                                 \
		::godot_lite::report_error("Index " #m_index " out of bounds (" #m_size ")", "", __FILE__, __LINE__); \
		return;                                                                   \
	} else                                                                        \
		((void)0)
#define ERR_FAIL_INDEX_MSG(m_index, m_size, m_msg)                                \
	if ((m_index) < 0 || (m_index) >= (m_size)) {                                 \
		::godot_lite::report_error("Index " #m_index " out of bounds (" #m_size ")", GL_MSG_STR(m_msg), __FILE__, __LINE__); \
		return;                                                                   \
	} else                                                                        \
		((void)0)
#define ERR_FAIL_INDEX_V(m_index, m_size, m_retval)                               \
	if ((m_index) < 0 || (m_index) >= (m_size)) {                                 \
		::godot_lite::report_error("Index " #m_index " out of bounds (" #m_size ")", "", __FILE__, __LINE__); \
		return m_retval;                                                          \
	} else                                                                        \
		((void)0)
#define ERR_FAIL_NULL(m_param)                                                    \
	if ((m_param) == nullptr) {                                                   \
		::godot_lite::report_error("Parameter \"" #m_param "\" is null", "", __FILE__, __LINE__); \
		return;                                                                   \
	} else                                                                        \
		((void)0)
#define ERR_FAIL_NULL_V(m_param, m_retval)                                        \
	if ((m_param) == nullptr) {                                                   \
		::godot_lite::report_error("Parameter \"" #m_param "\" is null", "", __FILE__, __LINE__); \
		return m_retval;                                                          \
	} else                                                                        \
		((void)0)
#define ERR_PRINT(m_msg) ::godot_lite::report_error("ERR_PRINT", GL_MSG_STR(m_msg), __FILE__, __LINE__)
#define WARN_PRINT(m_msg) ((void)0)
#define print_verbose(m_text) ((void)0)

/* ---- Math (core/math/math_funcs.h); float and double overloads chosen by argument type ----------- */
namespace Math {
constexpr double PI = 3.1415926535897932384626433833;
constexpr double TAU = 6.2831853071795864769252867666;
_ALWAYS_INLINE_ double sqrt(double p_x) { return ::sqrt(p_x); }
_ALWAYS_INLINE_ float sqrt(float p_x) { return ::sqrtf(p_x); }
_ALWAYS_INLINE_ double sin(double p_x) { return ::sin(p_x); }
_ALWAYS_INLINE_ float sin(float p_x) { return ::sinf(p_x); }
_ALWAYS_INLINE_ double cos(double p_x) { return ::cos(p_x); }
_ALWAYS_INLINE_ float cos(float p_x) { return ::cosf(p_x); }
_ALWAYS_INLINE_ double log(double p_x) { return ::log(p_x); }
_ALWAYS_INLINE_ float log(float p_x) { return ::logf(p_x); }
_ALWAYS_INLINE_ double log2(double p_x) { return ::log2(p_x); }
_ALWAYS_INLINE_ float log2(float p_x) { return ::log2f(p_x); }
_ALWAYS_INLINE_ double exp(double p_x) { return ::exp(p_x); }
_ALWAYS_INLINE_ float exp(float p_x) { return ::expf(p_x); }
_ALWAYS_INLINE_ double pow(double p_x, double p_y) { return ::pow(p_x, p_y); }
_ALWAYS_INLINE_ float pow(float p_x, float p_y) { return ::powf(p_x, p_y); }
/* upstream clamps the argument so that rounding just outside [-1, 1] does not produce NaN */
_ALWAYS_INLINE_ double acos(double p_x) { return p_x < -1 ? PI : (p_x > 1 ? 0 : ::acos(p_x)); }
_ALWAYS_INLINE_ float acos(float p_x) { return p_x < -1 ? (float)PI : (p_x > 1 ? 0 : ::acosf(p_x)); }
_ALWAYS_INLINE_ double abs(double g) { return ::fabs(g); }
_ALWAYS_INLINE_ float abs(float g) { return ::fabsf(g); }
_ALWAYS_INLINE_ int abs(int g) { return g > 0 ? g : -g; }
_ALWAYS_INLINE_ bool is_nan(double p_val) { return std::isnan(p_val); }
_ALWAYS_INLINE_ bool is_nan(float p_val) { return std::isnan(p_val); }
_ALWAYS_INLINE_ double rad_to_deg(double p_y) { return p_y * (180.0 / PI); }
_ALWAYS_INLINE_ float rad_to_deg(float p_y) { return p_y * (float)(180.0 / PI); }
_ALWAYS_INLINE_ double lerp(double p_from, double p_to, double p_weight) { return p_from + (p_to - p_from) * p_weight; }
_ALWAYS_INLINE_ float lerp(float p_from, float p_to, float p_weight) { return p_from + (p_to - p_from) * p_weight; }
_ALWAYS_INLINE_ double linear_to_db(double p_linear) { return Math::log(p_linear) * 8.6858896380650365530225783783321; }
_ALWAYS_INLINE_ float linear_to_db(float p_linear) { return Math::log(p_linear) * (float)8.6858896380650365530225783783321; }
_ALWAYS_INLINE_ double db_to_linear(double p_db) { return Math::exp(p_db * 0.11512925464970228420089957273422); }
_ALWAYS_INLINE_ float db_to_linear(float p_db) { return Math::exp(p_db * (float)0.11512925464970228420089957273422); }
} // namespace Math

/* ---- String / StringName (just enough) ------------------------------------------------------------ */
class StringName;
class String {
public:
	std::string s;
	String() {}
	String(const char *p) :
			s(p ? p : "") {}
	String(const std::string &p) :
			s(p) {}
	String(const StringName &p_name);
	String operator+(const String &o) const { return String(s + o.s); }
	String &operator+=(const String &o) {
		s += o.s;
		return *this;
	}
	bool operator==(const String &o) const { return s == o.s; }
	bool operator!=(const String &o) const { return s != o.s; }
	bool operator==(const char *o) const { return s == o; }
	bool is_empty() const { return s.empty(); }
};
class StringName {
public:
	std::string s;
	StringName() {}
	StringName(const char *p) :
			s(p ? p : "") {}
	StringName(const String &p) :
			s(p.s) {}
	bool operator==(const StringName &o) const { return s == o.s; }
	bool operator!=(const StringName &o) const { return s != o.s; }
	bool operator==(const char *o) const { return s == o; }
	uint32_t hash() const { return (uint32_t)std::hash<std::string>()(s); }
};
inline String::String(const StringName &p_name) :
		s(p_name.s) {}
#define SNAME(m_arg) StringName(m_arg)
#define SceneStringName(m_name) StringName(#m_name)
namespace godot_lite {
inline const char *to_cstr(const char *m) { return m; }
inline const char *to_cstr(const String &m) { return m.s.c_str(); }
} // namespace godot_lite
template <typename... VarArgs>
String vformat(const String &p_text, const VarArgs... p_args) {
	return p_text; /* arguments are only ever used for log text */
}

/* ---- Vector2 / Vector3 / Basis / Transform3D (core/math), real_t = float ------------------------------ */
struct Vector2 {
	real_t x = 0, y = 0;
	Vector2() {}
	Vector2(real_t p_x, real_t p_y) :
			x(p_x), y(p_y) {}
	real_t &operator[](int p_idx) { return p_idx == 0 ? x : y; }
	const real_t &operator[](int p_idx) const { return p_idx == 0 ? x : y; }
	Vector2 operator*(real_t p_rvalue) const { return Vector2(x * p_rvalue, y * p_rvalue); }
	void operator*=(real_t p_rvalue) {
		x *= p_rvalue;
		y *= p_rvalue;
	}
	Vector2 operator*(const Vector2 &p_v1) const { return Vector2(x * p_v1.x, y * p_v1.y); }
	Vector2 operator+(const Vector2 &p_v) const { return Vector2(x + p_v.x, y + p_v.y); }
	Vector2 operator-(const Vector2 &p_v) const { return Vector2(x - p_v.x, y - p_v.y); }
	bool operator==(const Vector2 &p_v) const { return x == p_v.x && y == p_v.y; }
	Vector2 lerp(const Vector2 &p_to, real_t p_weight) const {
		Vector2 res = *this;
		res.x = Math::lerp(res.x, p_to.x, p_weight);
		res.y = Math::lerp(res.y, p_to.y, p_weight);
		return res;
	}
};
_FORCE_INLINE_ Vector2 operator*(float p_scalar, const Vector2 &p_vec) { return p_vec * p_scalar; }
_FORCE_INLINE_ Vector2 operator*(double p_scalar, const Vector2 &p_vec) { return p_vec * p_scalar; }

struct Vector3 {
	real_t x = 0, y = 0, z = 0;
	Vector3() {}
	Vector3(real_t p_x, real_t p_y, real_t p_z) :
			x(p_x), y(p_y), z(p_z) {}
	real_t &operator[](int p_axis) { return (&x)[p_axis]; }
	const real_t &operator[](int p_axis) const { return (&x)[p_axis]; }
	real_t dot(const Vector3 &p_with) const { return x * p_with.x + y * p_with.y + z * p_with.z; }
	real_t length_squared() const {
		real_t x2 = x * x;
		real_t y2 = y * y;
		real_t z2 = z * z;
		return x2 + y2 + z2;
	}
	real_t length() const {
		real_t x2 = x * x;
		real_t y2 = y * y;
		real_t z2 = z * z;
		return Math::sqrt(x2 + y2 + z2);
	}
	void normalize() {
		real_t lengthsq = length_squared();
		if (lengthsq == 0) {
			x = y = z = 0;
		} else {
			real_t length = Math::sqrt(lengthsq);
			x /= length;
			y /= length;
			z /= length;
		}
	}
	Vector3 normalized() const {
		Vector3 v = *this;
		v.normalize();
		return v;
	}
	Vector3 operator+(const Vector3 &p_v) const { return Vector3(x + p_v.x, y + p_v.y, z + p_v.z); }
	Vector3 operator-(const Vector3 &p_v) const { return Vector3(x - p_v.x, y - p_v.y, z - p_v.z); }
	Vector3 operator-() const { return Vector3(-x, -y, -z); }
	Vector3 operator*(real_t p_scalar) const { return Vector3(x * p_scalar, y * p_scalar, z * p_scalar); }
	bool operator==(const Vector3 &p_v) const { return x == p_v.x && y == p_v.y && z == p_v.z; }
	bool operator!=(const Vector3 &p_v) const { return x != p_v.x || y != p_v.y || z != p_v.z; }
};

struct Basis {
	Vector3 rows[3] = { Vector3(1, 0, 0), Vector3(0, 1, 0), Vector3(0, 0, 1) };
	const Vector3 &operator[](int p_row) const { return rows[p_row]; }
	Vector3 &operator[](int p_row) { return rows[p_row]; }
	Vector3 get_column(int p_index) const { return Vector3(rows[0][p_index], rows[1][p_index], rows[2][p_index]); }
	void set_column(int p_index, const Vector3 &p_value) {
		rows[0][p_index] = p_value.x;
		rows[1][p_index] = p_value.y;
		rows[2][p_index] = p_value.z;
	}
	void set(real_t p_xx, real_t p_xy, real_t p_xz, real_t p_yx, real_t p_yy, real_t p_yz, real_t p_zx, real_t p_zy, real_t p_zz) {
		rows[0] = Vector3(p_xx, p_xy, p_xz);
		rows[1] = Vector3(p_yx, p_yy, p_yz);
		rows[2] = Vector3(p_zx, p_zy, p_zz);
	}
	void orthonormalize() { /* Gram-Schmidt */
		Vector3 x = get_column(0);
		Vector3 y = get_column(1);
		Vector3 z = get_column(2);
		x.normalize();
		y = (y - x * (x.dot(y)));
		y.normalize();
		z = (z - x * (x.dot(z)) - y * (y.dot(z)));
		z.normalize();
		set_column(0, x);
		set_column(1, y);
		set_column(2, z);
	}
	Basis orthonormalized() const {
		Basis c = *this;
		c.orthonormalize();
		return c;
	}
	void invert() {
#define GL_COFAC(row1, col1, row2, col2) (rows[row1][col1] * rows[row2][col2] - rows[row1][col2] * rows[row2][col1])
		real_t co[3] = { GL_COFAC(1, 1, 2, 2), GL_COFAC(1, 2, 2, 0), GL_COFAC(1, 0, 2, 1) };
		real_t det = rows[0][0] * co[0] + rows[0][1] * co[1] + rows[0][2] * co[2];
		ERR_FAIL_COND(det == 0);
		real_t s = 1.0f / det;
		set(co[0] * s, GL_COFAC(0, 2, 2, 1) * s, GL_COFAC(0, 1, 1, 2) * s,
				co[1] * s, GL_COFAC(0, 0, 2, 2) * s, GL_COFAC(0, 2, 1, 0) * s,
				co[2] * s, GL_COFAC(0, 1, 2, 0) * s, GL_COFAC(0, 0, 1, 1) * s);
#undef GL_COFAC
	}
	Vector3 xform(const Vector3 &p_vector) const {
		return Vector3(rows[0].dot(p_vector), rows[1].dot(p_vector), rows[2].dot(p_vector));
	}
	Vector3 xform_inv(const Vector3 &p_vector) const {
		return Vector3(
				(rows[0][0] * p_vector.x) + (rows[1][0] * p_vector.y) + (rows[2][0] * p_vector.z),
				(rows[0][1] * p_vector.x) + (rows[1][1] * p_vector.y) + (rows[2][1] * p_vector.z),
				(rows[0][2] * p_vector.x) + (rows[1][2] * p_vector.y) + (rows[2][2] * p_vector.z));
	}
};

struct Transform3D {
	Basis basis;
	Vector3 origin;
	void orthonormalize() { basis.orthonormalize(); }
	Transform3D orthonormalized() const {
		Transform3D c = *this;
		c.orthonormalize();
		return c;
	}
	void affine_invert() {
		basis.invert();
		origin = basis.xform(-origin);
	}
	Transform3D affine_inverse() const {
		Transform3D ret = *this;
		ret.affine_invert();
		return ret;
	}
	Vector3 xform(const Vector3 &p_vector) const {
		return Vector3(
				basis[0].dot(p_vector) + origin.x,
				basis[1].dot(p_vector) + origin.y,
				basis[2].dot(p_vector) + origin.z);
	}
};

/* ---- containers (core/templates): value semantics like upstream's copy-on-write ----------------------- */
template <typename T>
class Vector {
public:
	/* `write` must be the first member: its operator[] finds the owning Vector at its own address
	 * (upstream's VectorWriteProxy does the same with a computed offset). */
	struct WriteProxy {
		T &operator[](int64_t p_index) { return reinterpret_cast<Vector<T> *>(this)->data[(size_t)p_index]; }
	} write;

private:
	std::vector<T> data;

public:
	Vector() {}
	Vector(std::initializer_list<T> p_init) :
			data(p_init) {}
	Vector(const Vector &p_from) :
			write(), data(p_from.data) {}
	Vector &operator=(const Vector &p_from) {
		data = p_from.data;
		return *this;
	}
	int64_t size() const { return (int64_t)data.size(); }
	bool is_empty() const { return data.empty(); }
	int resize(int64_t p_size) {
		data.resize((size_t)p_size);
		return 0;
	}
	void fill(T p_elem) { std::fill(data.begin(), data.end(), p_elem); }
	void clear() { data.clear(); }
	bool push_back(T p_elem) {
		data.push_back(p_elem);
		return false;
	}
	void remove_at(int64_t p_index) { data.erase(data.begin() + p_index); }
	bool erase(const T &p_val) {
		for (size_t i = 0; i < data.size(); i++) {
			if (data[i] == p_val) {
				data.erase(data.begin() + i);
				return true;
			}
		}
		return false;
	}
	const T &operator[](int64_t p_index) const { return data[(size_t)p_index]; }
	const T *ptr() const { return data.data(); }
	T *ptrw() { return data.data(); }
	typename std::vector<T>::iterator begin() { return data.begin(); }
	typename std::vector<T>::iterator end() { return data.end(); }
	typename std::vector<T>::const_iterator begin() const { return data.begin(); }
	typename std::vector<T>::const_iterator end() const { return data.end(); }
};

template <typename T>
class LocalVector {
	std::vector<T> data;

public:
	uint32_t size() const { return (uint32_t)data.size(); }
	void clear() { data.clear(); }
	void push_back(T p_elem) { data.push_back(p_elem); }
	T &operator[](uint32_t i) { return data[i]; }
	const T &operator[](uint32_t i) const { return data[i]; }
	typename std::vector<T>::iterator begin() { return data.begin(); }
	typename std::vector<T>::iterator end() { return data.end(); }
	typename std::vector<T>::const_iterator begin() const { return data.begin(); }
	typename std::vector<T>::const_iterator end() const { return data.end(); }
};

template <typename T>
class List {
	std::list<T> data;

public:
	void push_back(const T &v) { data.push_back(v); }
	int size() const { return (int)data.size(); }
	typename std::list<T>::iterator begin() { return data.begin(); }
	typename std::list<T>::iterator end() { return data.end(); }
	typename std::list<T>::const_iterator begin() const { return data.begin(); }
	typename std::list<T>::const_iterator end() const { return data.end(); }
};

template <typename K, typename V>
struct KeyValue {
	const K key;
	V value;
	KeyValue(const K &k, const V &v) :
			key(k), value(v) {}
};

/* upstream HashMap iterates in insertion order */
template <typename K, typename V>
class HashMap {
	std::list<KeyValue<K, V>> items;

public:
	HashMap() {}
	HashMap(const HashMap &o) :
			items(o.items) {}
	HashMap &operator=(const HashMap &o) {
		items.clear();
		for (const KeyValue<K, V> &kv : o.items) {
			items.emplace_back(kv.key, kv.value);
		}
		return *this;
	}
	V *getptr(const K &k) {
		for (KeyValue<K, V> &kv : items) {
			if (kv.key == k) {
				return &kv.value;
			}
		}
		return nullptr;
	}
	const V *getptr(const K &k) const {
		for (const KeyValue<K, V> &kv : items) {
			if (kv.key == k) {
				return &kv.value;
			}
		}
		return nullptr;
	}
	bool has(const K &k) const { return getptr(k) != nullptr; }
	void insert(const K &k, const V &v) {
		if (V *p = getptr(k)) {
			*p = v;
		} else {
			items.emplace_back(k, v);
		}
	}
	V &operator[](const K &k) {
		if (V *p = getptr(k)) {
			return *p;
		}
		items.emplace_back(k, V());
		return items.back().value;
	}
	uint32_t size() const { return (uint32_t)items.size(); }
	typename std::list<KeyValue<K, V>>::iterator begin() { return items.begin(); }
	typename std::list<KeyValue<K, V>>::iterator end() { return items.end(); }
	typename std::list<KeyValue<K, V>>::const_iterator begin() const { return items.begin(); }
	typename std::list<KeyValue<K, V>>::const_iterator end() const { return items.end(); }
};

/* upstream HashSet iterates in insertion order */
template <typename T>
class HashSet {
	std::vector<T> items;

public:
	void insert(const T &v) {
		if (std::find(items.begin(), items.end(), v) == items.end()) {
			items.push_back(v);
		}
	}
	void clear() { items.clear(); }
	typename std::vector<T>::const_iterator begin() const { return items.begin(); }
	typename std::vector<T>::const_iterator end() const { return items.end(); }
};

/* ---- thread primitives (core/os/mutex.h, core/templates/safe_refcount.h, safe_list.h) -------------------- */
class Mutex {
	mutable std::recursive_mutex m;

public:
	void lock() const { m.lock(); }
	void unlock() const { m.unlock(); }
};
class SafeFlag {
	std::atomic<bool> flag{ false };

public:
	bool is_set() const { return flag.load(); }
	void set() { flag.store(true); }
	void clear() { flag.store(false); }
	void set_to(bool v) { flag.store(v); }
};
template <typename T>
class SafeNumeric {
	std::atomic<T> value;

public:
	explicit SafeNumeric(T p_value = static_cast<T>(0)) { value.store(p_value); }
	void set(T p_value) { value.store(p_value); }
	T get() const { return value.load(); }
};

/* upstream SafeList: lock-free, inserts at the HEAD, erase() defers the deleter until maybe_cleanup()
 * and leaves the erased node's `next` intact so an iteration in flight carries on. */
template <typename T>
class SafeList {
	struct Node {
		T val;
		Node *next = nullptr;
		std::function<void(T)> deletion_fn;
		bool erased = false;
	};
	Node *head = nullptr;
	std::vector<Node *> graveyard;

public:
	class Iterator {
		Node *cursor;

	public:
		Iterator(Node *p) :
				cursor(p) {
			skip();
		}
		void skip() {
			while (cursor && cursor->erased) {
				cursor = cursor->next;
			}
		}
		T &operator*() { return cursor->val; }
		Iterator &operator++() {
			cursor = cursor->next;
			skip();
			return *this;
		}
		bool operator!=(const Iterator &o) const { return cursor != o.cursor; }
	};
	void insert(T p_value) {
		Node *n = new Node();
		n->val = p_value;
		n->next = head;
		head = n;
	}
	void erase(T p_value, std::function<void(T)> p_deletion_fn) {
		Node **link = &head;
		for (Node *n = head; n; n = n->next) {
			if (!n->erased && n->val == p_value) {
				*link = n->next; /* unlink; n->next stays for iterators standing on n */
				n->erased = true;
				n->deletion_fn = p_deletion_fn;
				graveyard.push_back(n);
				return;
			}
			link = &n->next;
		}
	}
	void erase(T p_value) {
		erase(p_value, [](T) {});
	}
	Iterator begin() { return Iterator(head); }
	Iterator end() { return Iterator(nullptr); }
	bool maybe_cleanup() {
		for (Node *n : graveyard) {
			if (n->deletion_fn) {
				n->deletion_fn(n->val);
			}
			delete n;
		}
		graveyard.clear();
		return true;
	}
	~SafeList() {
		maybe_cleanup();
		while (head) {
			Node *n = head;
			head = n->next;
			delete n;
		}
	}
};

/* ---- Object / RefCounted / Ref (core/object) ------------------------------------------------------------ */
class Callable {
public:
	Callable unbind(int) const { return *this; }
};
template <typename T, typename M>
Callable callable_mp(T *, M) {
	return Callable();
}
enum Error { OK = 0, FAILED = 1 };

class Object {
public:
	virtual ~Object() {}
	template <typename T>
	static T *cast_to(Object *p_object) { return dynamic_cast<T *>(p_object); }
	template <typename T>
	static const T *cast_to(const Object *p_object) { return dynamic_cast<const T *>(p_object); }
	void notify_property_list_changed() {}
	Error connect(const StringName &, const Callable &, uint32_t = 0) { return OK; }
	void disconnect(const StringName &, const Callable &) {}
	int emitted_signals = 0;
	template <typename... A>
	Error emit_signal(const StringName &, A...) {
		emitted_signals++;
		return OK;
	}
	/* notification plumbing: GDCLASS overrides _notificationv the way upstream does */
	void _notification(int) {}
	virtual void _notificationv(int) {}
	void notification(int p_what) { _notificationv(p_what); }
	static void _bind_methods() {}
};

#define GDCLASS(m_class, m_inherits)                                                                          \
public:                                                                                                       \
	typedef m_class self_type;                                                                                \
	typedef m_inherits super_type;                                                                            \
	static const char *get_class_static() { return #m_class; }                                                \
	virtual void _notificationv(int p_what) override {                                                        \
		m_inherits::_notificationv(p_what);                                                                   \
		if constexpr (!std::is_same_v<decltype(&m_class::_notification), decltype(&m_inherits::_notification)>) { \
			m_class::_notification(p_what);                                                                   \
		}                                                                                                     \
	}                                                                                                         \
                                                                                                              \
private:

class RefCounted : public Object {
	std::atomic<int> refcount{ 0 };

public:
	void reference() { refcount.fetch_add(1); }
	bool unreference() { return refcount.fetch_sub(1) == 1; }
	int get_reference_count() const { return refcount.load(); }
};

template <typename T>
class Ref {
	T *reference = nullptr;
	void ref_pointer(T *p) {
		if (p == reference) {
			return;
		}
		T *old = reference;
		reference = p;
		if (reference) {
			reference->reference();
		}
		if (old && old->unreference()) {
			delete old;
		}
	}

public:
	Ref() {}
	Ref(T *p_reference) { ref_pointer(p_reference); }
	Ref(const Ref &p_from) { ref_pointer(p_from.reference); }
	template <typename T_Other>
	Ref(const Ref<T_Other> &p_from) {
		ref_pointer(dynamic_cast<T *>(static_cast<Object *>(p_from.ptr())));
	}
	~Ref() { unref(); }
	Ref &operator=(const Ref &p_from) {
		ref_pointer(p_from.reference);
		return *this;
	}
	template <typename T_Other>
	Ref &operator=(const Ref<T_Other> &p_from) {
		ref_pointer(dynamic_cast<T *>(static_cast<Object *>(p_from.ptr())));
		return *this;
	}
	bool operator==(const T *p_ptr) const { return reference == p_ptr; }
	bool operator==(const Ref &p_r) const { return reference == p_r.reference; }
	bool operator!=(const Ref &p_r) const { return reference != p_r.reference; }
	T *operator->() const { return reference; }
	T *operator*() const { return reference; }
	T *ptr() const { return reference; }
	bool is_valid() const { return reference != nullptr; }
	bool is_null() const { return reference == nullptr; }
	void unref() { ref_pointer(nullptr); }
	void instantiate() { ref_pointer(new T()); }
};

class Resource : public RefCounted {
	GDCLASS(Resource, RefCounted);

public:
	virtual Ref<Resource> duplicate(bool p_subresources = false) const { return Ref<Resource>(); }
};

#define memnew(m_class) new m_class
template <typename T>
void memdelete(T *p) {
	delete p;
}

/* ---- Variant / Dictionary / TypedArray: only the shapes the module touches ------------------------------- */
class Variant {
public:
	enum Type { NIL, BOOL, INT, FLOAT, STRING, STRING_NAME, OBJECT, ARRAY, PACKED_VECTOR2_ARRAY };

private:
	Type type = NIL;
	double num = 0;
	std::string str;
	Vector<Vector2> v2;

public:
	Variant() {}
	Variant(bool v) :
			type(BOOL), num(v) {}
	Variant(int v) :
			type(INT), num(v) {}
	Variant(int64_t v) :
			type(INT), num((double)v) {}
	Variant(float v) :
			type(FLOAT), num(v) {}
	Variant(double v) :
			type(FLOAT), num(v) {}
	Variant(const char *v) :
			type(STRING), str(v) {}
	Variant(const String &v) :
			type(STRING), str(v.s) {}
	Variant(const StringName &v) :
			type(STRING_NAME), str(v.s) {}
	Variant(const Vector<Vector2> &v) :
			type(PACKED_VECTOR2_ARRAY), v2(v) {}
	Type get_type() const { return type; }
	operator StringName() const { return StringName(str.c_str()); }
	operator String() const { return String(str); }
	operator Vector<Vector2>() const { return v2; }
	operator float() const { return (float)num; }
	operator bool() const { return num != 0; }
	bool operator==(const Variant &o) const {
		bool s1 = type == STRING || type == STRING_NAME, s2 = o.type == STRING || o.type == STRING_NAME;
		if (s1 && s2) {
			return str == o.str;
		}
		if (type != o.type) {
			return false;
		}
		if (type == PACKED_VECTOR2_ARRAY) {
			if (v2.size() != o.v2.size()) {
				return false;
			}
			for (int64_t i = 0; i < v2.size(); i++) {
				if (!(v2[i] == o.v2[i])) {
					return false;
				}
			}
			return true;
		}
		return num == o.num;
	}
};

/* upstream Dictionary: reference-counted shared storage, insertion-ordered keys */
class Dictionary {
	struct Storage {
		std::list<std::pair<Variant, Variant>> items;
	};
	std::shared_ptr<Storage> p = std::make_shared<Storage>();

public:
	Variant &operator[](const Variant &p_key) {
		for (auto &kv : p->items) {
			if (kv.first == p_key) {
				return kv.second;
			}
		}
		p->items.emplace_back(p_key, Variant());
		return p->items.back().second;
	}
	Variant get_valid(const Variant &p_key) const {
		for (auto &kv : p->items) {
			if (kv.first == p_key) {
				return kv.second;
			}
		}
		return Variant();
	}
	LocalVector<Variant> get_key_list() const {
		LocalVector<Variant> keys;
		for (auto &kv : p->items) {
			keys.push_back(kv.first);
		}
		return keys;
	}
	int size() const { return (int)p->items.size(); }
};

template <typename T>
class TypedArray {
	std::vector<Ref<T>> items;

public:
	int size() const { return (int)items.size(); }
	void push_back(const Ref<T> &v) { items.push_back(v); }
	Ref<T> operator[](int i) const { return items[(size_t)i]; }
};

/* ---- ClassDB / property registration: accepted and ignored ------------------------------------------------ */
enum PropertyHint { PROPERTY_HINT_NONE };
enum PropertyUsageFlags { PROPERTY_USAGE_NONE = 0, PROPERTY_USAGE_STORAGE = 2, PROPERTY_USAGE_EDITOR = 4 };
struct PropertyInfo {
	String name;
	String hint_string;
	uint32_t usage = 6;
};
struct ClassDB {
	template <typename... A>
	static void bind_method(A &&...) {}
};
#define D_METHOD(...) 0
#define DEFVAL(m_defval) 0
#define ADD_PROPERTY(...) ((void)0)
#define ADD_GROUP(...) ((void)0)
#define ADD_SIGNAL(...) ((void)0)
#define BIND_ENUM_CONSTANT(m_constant) ((void)0)
#define VARIANT_ENUM_CAST(m_enum)
#define GDREGISTER_CLASS(m_class) ::godot_lite::registered_classes().push_back(#m_class)
#define GDREGISTER_VIRTUAL_CLASS(m_class) ::godot_lite::registered_classes().push_back(#m_class)
#define GDREGISTER_ABSTRACT_CLASS(m_class) ::godot_lite::registered_classes().push_back(#m_class)
namespace godot_lite {
inline std::vector<std::string> &registered_classes() {
	static std::vector<std::string> v;
	return v;
}
} // namespace godot_lite
enum ModuleInitializationLevel {
	MODULE_INITIALIZATION_LEVEL_CORE,
	MODULE_INITIALIZATION_LEVEL_SERVERS,
	MODULE_INITIALIZATION_LEVEL_SCENE,
	MODULE_INITIALIZATION_LEVEL_EDITOR
};

/* ---- script virtuals (core/object/gdvirtual.gen.inc): a std::function hook plays the script ---------------- */
template <typename T>
struct GDExtensionPtr {
	T *data = nullptr;
	GDExtensionPtr() {}
	GDExtensionPtr(T *p) :
			data(p) {}
	operator T *() const { return data; }
};
template <typename T>
struct GDExtensionConstPtr {
	const T *data = nullptr;
	GDExtensionConstPtr() {}
	GDExtensionConstPtr(const T *p) :
			data(p) {}
	operator const T *() const { return data; }
};
#define GL_VHOOK(m_name) _gdvirtual_##m_name##_hook
#define GL_VCALL(m_name) _gdvirtual_##m_name##_call
#define GDVIRTUAL0(m_name)                        \
	mutable std::function<bool()> GL_VHOOK(m_name); \
	bool GL_VCALL(m_name)() { return GL_VHOOK(m_name) ? GL_VHOOK(m_name)() : false; }
#define GDVIRTUAL0R(m_ret, m_name)                       \
	mutable std::function<bool(m_ret &)> GL_VHOOK(m_name); \
	bool GL_VCALL(m_name)(m_ret & r_ret) { return GL_VHOOK(m_name) ? GL_VHOOK(m_name)(r_ret) : false; }
#define GDVIRTUAL0RC(m_ret, m_name)                      \
	mutable std::function<bool(m_ret &)> GL_VHOOK(m_name); \
	bool GL_VCALL(m_name)(m_ret & r_ret) const { return GL_VHOOK(m_name) ? GL_VHOOK(m_name)(r_ret) : false; }
#define GDVIRTUAL0R_REQUIRED(m_ret, m_name) GDVIRTUAL0R(m_ret, m_name)
#define GDVIRTUAL2(m_name, A1, A2)                      \
	mutable std::function<bool(A1, A2)> GL_VHOOK(m_name); \
	bool GL_VCALL(m_name)(A1 a1, A2 a2) { return GL_VHOOK(m_name) ? GL_VHOOK(m_name)(a1, a2) : false; }
#define GDVIRTUAL5(m_name, A1, A2, A3, A4, A5)                      \
	mutable std::function<bool(A1, A2, A3, A4, A5)> GL_VHOOK(m_name); \
	bool GL_VCALL(m_name)(A1 a1, A2 a2, A3 a3, A4 a4, A5 a5) { return GL_VHOOK(m_name) ? GL_VHOOK(m_name)(a1, a2, a3, a4, a5) : false; }
#define GDVIRTUAL6(m_name, A1, A2, A3, A4, A5, A6)                      \
	mutable std::function<bool(A1, A2, A3, A4, A5, A6)> GL_VHOOK(m_name); \
	bool GL_VCALL(m_name)(A1 a1, A2 a2, A3 a3, A4 a4, A5 a5, A6 a6) { return GL_VHOOK(m_name) ? GL_VHOOK(m_name)(a1, a2, a3, a4, a5, a6) : false; }
#define GDVIRTUAL_CALL(m_name, ...) GL_VCALL(m_name)(__VA_ARGS__)
#define GDVIRTUAL_BIND(m_name, ...) ((void)0)

/* ---- project settings / engine ---------------------------------------------------------------------------- */
namespace godot_lite {
inline float &global_3d_panning_strength() {
	static float v = 0.5f; /* audio/general/3d_panning_strength default */
	return v;
}
inline float project_setting_float(const char *) { return global_3d_panning_strength(); }
} // namespace godot_lite
#define GLOBAL_GET_CACHED(m_type, m_setting_name) ((m_type)::godot_lite::project_setting_float(m_setting_name))
class Engine {
public:
	static Engine *get_singleton() {
		static Engine e;
		return &e;
	}
	bool is_editor_hint() const { return false; }
};
