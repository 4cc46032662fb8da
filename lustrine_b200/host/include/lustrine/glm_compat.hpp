// glm_compat.hpp — the handful of glm 0.9.9.8 types the Lustrine API exposes (vec3, vec4).
//
// The reference's public headers pass glm::vec3 / glm::vec4 by value.  glm is a third-party
// header library that is not vendored here; define LUSTRINE_USE_SYSTEM_GLM to build against a real
// glm instead.  The templates below have the same names, template parameters and member layout
// as glm's (vec<L, T, Q> with Q = defaultp = packed_highp), so symbols mangle identically and a
// caller compiled against real glm links against this library.
#pragma once

#ifdef LUSTRINE_USE_SYSTEM_GLM
#include <glm/glm.hpp>
#else
#include <cmath>

namespace glm {

typedef int length_t;
enum qualifier { packed_highp, packed_mediump, packed_lowp, highp = packed_highp, mediump = packed_mediump, lowp = packed_lowp, packed = packed_highp, defaultp = highp };

template <length_t L, typename T, qualifier Q = defaultp> struct vec;

template <typename T, qualifier Q> struct vec<3, T, Q> {
    union { T x, r, s; };
    union { T y, g, t; };
    union { T z, b, p; };
    vec() = default;
    explicit vec(T v) : x(v), y(v), z(v) {}
    template <typename A, typename B, typename C> vec(A a, B b_, C c) : x((T)a), y((T)b_), z((T)c) {}
    T& operator[](int i) { return (&x)[i]; }
    const T& operator[](int i) const { return (&x)[i]; }
    vec& operator+=(const vec& o) { x += o.x; y += o.y; z += o.z; return *this; }
    vec& operator-=(const vec& o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
    vec& operator*=(T s_) { x *= s_; y *= s_; z *= s_; return *this; }
    vec& operator/=(T s_) { x /= s_; y /= s_; z /= s_; return *this; }
    static constexpr length_t length() { return 3; }
};

template <typename T, qualifier Q> struct vec<4, T, Q> {
    union { T x, r, s; };
    union { T y, g, t; };
    union { T z, b, p; };
    union { T w, a, q; };
    vec() = default;
    explicit vec(T v) : x(v), y(v), z(v), w(v) {}
    template <typename A, typename B, typename C, typename D> vec(A a_, B b_, C c, D d) : x((T)a_), y((T)b_), z((T)c), w((T)d) {}
    T& operator[](int i) { return (&x)[i]; }
    const T& operator[](int i) const { return (&x)[i]; }
};

typedef vec<3, float, defaultp> vec3;
typedef vec<4, float, defaultp> vec4;

template <typename T, qualifier Q> inline vec<3, T, Q> operator+(const vec<3, T, Q>& a, const vec<3, T, Q>& b) { return vec<3, T, Q>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <typename T, qualifier Q> inline vec<3, T, Q> operator-(const vec<3, T, Q>& a, const vec<3, T, Q>& b) { return vec<3, T, Q>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <typename T, qualifier Q> inline vec<3, T, Q> operator-(const vec<3, T, Q>& a) { return vec<3, T, Q>(-a.x, -a.y, -a.z); }
template <typename T, qualifier Q> inline vec<3, T, Q> operator*(const vec<3, T, Q>& a, T s) { return vec<3, T, Q>(a.x * s, a.y * s, a.z * s); }
template <typename T, qualifier Q> inline vec<3, T, Q> operator*(T s, const vec<3, T, Q>& a) { return vec<3, T, Q>(s * a.x, s * a.y, s * a.z); }
template <typename T, qualifier Q> inline vec<3, T, Q> operator/(const vec<3, T, Q>& a, T s) { return vec<3, T, Q>(a.x / s, a.y / s, a.z / s); }
template <typename T, qualifier Q> inline T dot(const vec<3, T, Q>& a, const vec<3, T, Q>& b) { T tx = a.x * b.x, ty = a.y * b.y, tz = a.z * b.z; return (tx + ty) + tz; }
template <typename T, qualifier Q> inline T length(const vec<3, T, Q>& a) { return std::sqrt(dot(a, a)); }
template <typename T> inline T clamp(T v, T lo, T hi) { return v < lo ? lo : (v > hi ? hi : v); }

}  // namespace glm
#endif
