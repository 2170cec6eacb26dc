// Minimal stand-in for the parts of GLM (0.9.9.8 scalar path) that the
// reference's chunk light + mesh translation units reach.
//
// TEST INFRASTRUCTURE ONLY: this header exists so that the reference's own
// sources can be compiled, unmodified, from /root/reference into
// oracle/_ref/libvxref.so (GLM is an un-vendored dependency of the reference,
// vcpkg.json:9, and is not installed in this image).  Nothing in the product
// (voxelengine-cpp_b200/) includes it.
//
// Arithmetic follows GLM's non-SIMD definitions (no GLM_FORCE_INTRINSICS is
// defined anywhere in the reference):
//   dot(a,b)      = a.x*b.x + a.y*b.y + a.z*b.z          (left to right)
//   normalize(v)  = v * (1 / sqrt(dot(v,v)))
//   cross(x,y)    = (x.y*y.z - y.y*x.z, x.z*y.x - y.z*x.x, x.x*y.y - y.x*x.y)
//   everything else component-wise; converting constructors are implicit and
//   truncate (GLM_FORCE_EXPLICIT_CTOR is not defined by the reference).
#pragma once

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <functional>
#include <type_traits>

namespace glm {

enum qualifier { defaultp };

template <int L, typename T, qualifier Q = defaultp>
struct vec;

template <typename T, qualifier Q>
struct vec<2, T, Q> {
    union { T x, r, s; };
    union { T y, g, t; };
    constexpr vec() : x(0), y(0) {}
    template <typename U, typename = std::enable_if_t<std::is_arithmetic_v<U>>>
    constexpr vec(U s) : x(static_cast<T>(s)), y(static_cast<T>(s)) {}
    template <typename A, typename B>
    constexpr vec(A a, B b) : x(static_cast<T>(a)), y(static_cast<T>(b)) {}
    template <typename U>
    constexpr vec(const vec<2, U, Q>& v)
        : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)) {}
    template <typename U>
    constexpr vec(const vec<3, U, Q>& v);
    template <typename U>
    constexpr vec(const vec<4, U, Q>& v);
    constexpr T& operator[](int i) { return i == 0 ? x : y; }
    constexpr const T& operator[](int i) const { return i == 0 ? x : y; }
    static constexpr int length() { return 2; }
};

template <typename T, qualifier Q>
struct vec<3, T, Q> {
    union { T x, r, s; };
    union { T y, g, t; };
    union { T z, b, p; };
    constexpr vec() : x(0), y(0), z(0) {}
    template <typename U, typename = std::enable_if_t<std::is_arithmetic_v<U>>>
    constexpr vec(U s)
        : x(static_cast<T>(s)), y(static_cast<T>(s)), z(static_cast<T>(s)) {}
    template <typename A, typename B, typename C>
    constexpr vec(A a, B b, C c)
        : x(static_cast<T>(a)), y(static_cast<T>(b)), z(static_cast<T>(c)) {}
    template <typename U>
    constexpr vec(const vec<3, U, Q>& v)
        : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)), z(static_cast<T>(v.z)) {}
    template <typename U, typename C>
    constexpr vec(const vec<2, U, Q>& v, C c)
        : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)), z(static_cast<T>(c)) {}
    template <typename U>
    constexpr vec(const vec<4, U, Q>& v);
    constexpr T& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    constexpr const T& operator[](int i) const {
        return i == 0 ? x : (i == 1 ? y : z);
    }
    static constexpr int length() { return 3; }
};

template <typename T, qualifier Q>
struct vec<4, T, Q> {
    union { T x, r, s; };
    union { T y, g, t; };
    union { T z, b, p; };
    union { T w, a, q; };
    constexpr vec() : x(0), y(0), z(0), w(0) {}
    template <typename U, typename = std::enable_if_t<std::is_arithmetic_v<U>>>
    constexpr vec(U s)
        : x(static_cast<T>(s)), y(static_cast<T>(s)), z(static_cast<T>(s)),
          w(static_cast<T>(s)) {}
    template <typename A, typename B, typename C, typename D>
    constexpr vec(A a, B b, C c, D d)
        : x(static_cast<T>(a)), y(static_cast<T>(b)), z(static_cast<T>(c)),
          w(static_cast<T>(d)) {}
    template <typename U>
    constexpr vec(const vec<4, U, Q>& v)
        : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)), z(static_cast<T>(v.z)),
          w(static_cast<T>(v.w)) {}
    template <typename U, typename D>
    constexpr vec(const vec<3, U, Q>& v, D d)
        : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)), z(static_cast<T>(v.z)),
          w(static_cast<T>(d)) {}
    template <typename U, typename C, typename D>
    constexpr vec(const vec<2, U, Q>& v, C c, D d)
        : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)), z(static_cast<T>(c)),
          w(static_cast<T>(d)) {}
    constexpr T& operator[](int i) {
        return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w));
    }
    constexpr const T& operator[](int i) const {
        return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w));
    }
    static constexpr int length() { return 4; }
};

template <typename T, qualifier Q>
template <typename U>
constexpr vec<2, T, Q>::vec(const vec<3, U, Q>& v)
    : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)) {}
template <typename T, qualifier Q>
template <typename U>
constexpr vec<2, T, Q>::vec(const vec<4, U, Q>& v)
    : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)) {}
template <typename T, qualifier Q>
template <typename U>
constexpr vec<3, T, Q>::vec(const vec<4, U, Q>& v)
    : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)), z(static_cast<T>(v.z)) {}

using vec2 = vec<2, float>;
using vec3 = vec<3, float>;
using vec4 = vec<4, float>;
using dvec2 = vec<2, double>;
using dvec3 = vec<3, double>;
using dvec4 = vec<4, double>;
using ivec2 = vec<2, int>;
using ivec3 = vec<3, int>;
using ivec4 = vec<4, int>;
using uvec2 = vec<2, unsigned>;
using uvec3 = vec<3, unsigned>;
using uvec4 = vec<4, unsigned>;
using u32vec2 = vec<2, uint32_t>;
using i8vec3 = vec<3, int8_t>;
using bvec3 = vec<3, bool>;
using highp_dvec2 = dvec2;
using highp_dvec3 = dvec3;
using highp_vec3 = vec3;

// ---- component-wise operator generation -----------------------------------
#define GLMSHIM_FOR2(OP) return vec<2, T, Q>(a.x OP b.x, a.y OP b.y)
#define GLMSHIM_FOR3(OP) return vec<3, T, Q>(a.x OP b.x, a.y OP b.y, a.z OP b.z)
#define GLMSHIM_FOR4(OP) \
    return vec<4, T, Q>(a.x OP b.x, a.y OP b.y, a.z OP b.z, a.w OP b.w)

#define GLMSHIM_BINOP(OP)                                                      \
    template <typename T, qualifier Q>                                         \
    constexpr vec<2, T, Q> operator OP(const vec<2, T, Q>& a,                  \
                                       const vec<2, T, Q>& b) {                \
        GLMSHIM_FOR2(OP);                                                      \
    }                                                                          \
    template <typename T, qualifier Q>                                         \
    constexpr vec<3, T, Q> operator OP(const vec<3, T, Q>& a,                  \
                                       const vec<3, T, Q>& b) {                \
        GLMSHIM_FOR3(OP);                                                      \
    }                                                                          \
    template <typename T, qualifier Q>                                         \
    constexpr vec<4, T, Q> operator OP(const vec<4, T, Q>& a,                  \
                                       const vec<4, T, Q>& b) {                \
        GLMSHIM_FOR4(OP);                                                      \
    }                                                                          \
    template <int L, typename T, qualifier Q>                                  \
    constexpr vec<L, T, Q> operator OP(const vec<L, T, Q>& a, T s) {           \
        return a OP vec<L, T, Q>(s);                                           \
    }                                                                          \
    template <int L, typename T, qualifier Q>                                  \
    constexpr vec<L, T, Q> operator OP(T s, const vec<L, T, Q>& b) {           \
        return vec<L, T, Q>(s) OP b;                                           \
    }                                                                          \
    template <int L, typename T, qualifier Q>                                  \
    constexpr vec<L, T, Q>& operator OP##=(vec<L, T, Q>& a,                    \
                                           const vec<L, T, Q>& b) {            \
        a = a OP b;                                                            \
        return a;                                                              \
    }                                                                          \
    template <int L, typename T, qualifier Q, typename U,                      \
              typename = std::enable_if_t<std::is_arithmetic_v<U>>>            \
    constexpr vec<L, T, Q>& operator OP##=(vec<L, T, Q>& a, U s) {             \
        a = a OP vec<L, T, Q>(static_cast<T>(s));                              \
        return a;                                                              \
    }                                                                          \
    template <int L, typename T, typename U, qualifier Q,                      \
              typename = std::enable_if_t<!std::is_same_v<T, U>>>              \
    constexpr vec<L, T, Q>& operator OP##=(vec<L, T, Q>& a,                    \
                                           const vec<L, U, Q>& b) {            \
        a = a OP vec<L, T, Q>(b);                                              \
        return a;                                                              \
    }

GLMSHIM_BINOP(+)
GLMSHIM_BINOP(-)
GLMSHIM_BINOP(*)
GLMSHIM_BINOP(/)
#undef GLMSHIM_BINOP

template <typename T, qualifier Q>
constexpr vec<2, T, Q> operator-(const vec<2, T, Q>& a) {
    return vec<2, T, Q>(-a.x, -a.y);
}
template <typename T, qualifier Q>
constexpr vec<3, T, Q> operator-(const vec<3, T, Q>& a) {
    return vec<3, T, Q>(-a.x, -a.y, -a.z);
}
template <typename T, qualifier Q>
constexpr vec<4, T, Q> operator-(const vec<4, T, Q>& a) {
    return vec<4, T, Q>(-a.x, -a.y, -a.z, -a.w);
}

template <typename T, qualifier Q>
constexpr bool operator==(const vec<2, T, Q>& a, const vec<2, T, Q>& b) {
    return a.x == b.x && a.y == b.y;
}
template <typename T, qualifier Q>
constexpr bool operator==(const vec<3, T, Q>& a, const vec<3, T, Q>& b) {
    return a.x == b.x && a.y == b.y && a.z == b.z;
}
template <typename T, qualifier Q>
constexpr bool operator==(const vec<4, T, Q>& a, const vec<4, T, Q>& b) {
    return a.x == b.x && a.y == b.y && a.z == b.z && a.w == b.w;
}
template <int L, typename T, qualifier Q>
constexpr bool operator!=(const vec<L, T, Q>& a, const vec<L, T, Q>& b) {
    return !(a == b);
}

// integer-only operators used by a few headers
template <typename T, qualifier Q>
constexpr vec<3, T, Q> operator%(const vec<3, T, Q>& a, T s) {
    return vec<3, T, Q>(a.x % s, a.y % s, a.z % s);
}

// ---- functions ------------------------------------------------------------
template <typename T, qualifier Q>
constexpr T dot(const vec<2, T, Q>& a, const vec<2, T, Q>& b) {
    vec<2, T, Q> tmp(a * b);
    return tmp.x + tmp.y;
}
template <typename T, qualifier Q>
constexpr T dot(const vec<3, T, Q>& a, const vec<3, T, Q>& b) {
    vec<3, T, Q> tmp(a * b);
    return tmp.x + tmp.y + tmp.z;
}
template <typename T, qualifier Q>
constexpr T dot(const vec<4, T, Q>& a, const vec<4, T, Q>& b) {
    vec<4, T, Q> tmp(a * b);
    return (tmp.x + tmp.y) + (tmp.z + tmp.w);
}

template <typename T>
inline T inversesqrt(T x) {
    return static_cast<T>(1) / std::sqrt(x);
}

template <int L, typename T, qualifier Q>
inline vec<L, T, Q> normalize(const vec<L, T, Q>& v) {
    return v * inversesqrt(dot(v, v));
}

template <int L, typename T, qualifier Q>
inline T length(const vec<L, T, Q>& v) {
    return std::sqrt(dot(v, v));
}
template <int L, typename T, qualifier Q>
inline T length2(const vec<L, T, Q>& v) {
    return dot(v, v);
}
template <int L, typename T, qualifier Q>
inline T distance(const vec<L, T, Q>& a, const vec<L, T, Q>& b) {
    return length(b - a);
}
template <int L, typename T, qualifier Q>
inline T distance2(const vec<L, T, Q>& a, const vec<L, T, Q>& b) {
    return length2(b - a);
}

template <typename T, qualifier Q>
constexpr vec<3, T, Q> cross(const vec<3, T, Q>& x, const vec<3, T, Q>& y) {
    return vec<3, T, Q>(
        x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y
    );
}

template <typename T, typename = std::enable_if_t<std::is_arithmetic_v<T>>>
constexpr T min(T a, T b) {
    return (b < a) ? b : a;
}
template <typename T, typename = std::enable_if_t<std::is_arithmetic_v<T>>>
constexpr T max(T a, T b) {
    return (a < b) ? b : a;
}
template <typename T, qualifier Q>
constexpr vec<2, T, Q> min(const vec<2, T, Q>& a, const vec<2, T, Q>& b) {
    return vec<2, T, Q>(min(a.x, b.x), min(a.y, b.y));
}
template <typename T, qualifier Q>
constexpr vec<3, T, Q> min(const vec<3, T, Q>& a, const vec<3, T, Q>& b) {
    return vec<3, T, Q>(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z));
}
template <typename T, qualifier Q>
constexpr vec<4, T, Q> min(const vec<4, T, Q>& a, const vec<4, T, Q>& b) {
    return vec<4, T, Q>(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z), min(a.w, b.w));
}
template <typename T, qualifier Q>
constexpr vec<2, T, Q> max(const vec<2, T, Q>& a, const vec<2, T, Q>& b) {
    return vec<2, T, Q>(max(a.x, b.x), max(a.y, b.y));
}
template <typename T, qualifier Q>
constexpr vec<3, T, Q> max(const vec<3, T, Q>& a, const vec<3, T, Q>& b) {
    return vec<3, T, Q>(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z));
}
template <typename T, qualifier Q>
constexpr vec<4, T, Q> max(const vec<4, T, Q>& a, const vec<4, T, Q>& b) {
    return vec<4, T, Q>(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z), max(a.w, b.w));
}

template <typename T, typename = std::enable_if_t<std::is_arithmetic_v<T>>>
inline T floor(T v) {
    return std::floor(v);
}
template <typename T, qualifier Q>
inline vec<2, T, Q> floor(const vec<2, T, Q>& v) {
    return vec<2, T, Q>(std::floor(v.x), std::floor(v.y));
}
template <typename T, qualifier Q>
inline vec<3, T, Q> floor(const vec<3, T, Q>& v) {
    return vec<3, T, Q>(std::floor(v.x), std::floor(v.y), std::floor(v.z));
}
template <typename T, typename = std::enable_if_t<std::is_arithmetic_v<T>>>
inline T abs(T v) {
    return v < 0 ? -v : v;
}
template <typename T, qualifier Q>
inline vec<3, T, Q> abs(const vec<3, T, Q>& v) {
    return vec<3, T, Q>(abs(v.x), abs(v.y), abs(v.z));
}
template <typename T, typename = std::enable_if_t<std::is_arithmetic_v<T>>>
constexpr T clamp(T v, T lo, T hi) {
    return min(max(v, lo), hi);
}

template <typename T, typename U,
          typename = std::enable_if_t<std::is_arithmetic_v<T>>>
constexpr T mix(T a, T b, U t) {
    return static_cast<T>(a * (static_cast<U>(1) - t) + b * t);
}
template <typename T, qualifier Q>
constexpr vec<3, T, Q> mix(
    const vec<3, T, Q>& a, const vec<3, T, Q>& b, const vec<3, T, Q>& t
) {
    return a * (vec<3, T, Q>(static_cast<T>(1)) - t) + b * t;
}
template <typename T, qualifier Q>
constexpr vec<3, T, Q> mix(const vec<3, T, Q>& a, const vec<3, T, Q>& b, T t) {
    return a * (static_cast<T>(1) - t) + b * t;
}

template <typename T>
constexpr T radians(T deg) {
    return deg * static_cast<T>(0.01745329251994329576923690768489);
}
template <typename T>
constexpr T degrees(T rad) {
    return rad * static_cast<T>(57.295779513082320876798154814105);
}
template <typename T>
constexpr T pi() {
    return static_cast<T>(3.14159265358979323846264338327950288);
}

// ---- matrices (storage + the few operators headers mention) ---------------
template <int C, int R, typename T, qualifier Q = defaultp>
struct mat {
    vec<R, T, Q> value[C];
    constexpr mat() : value{} {}
    constexpr explicit mat(T s) : value{} {
        for (int i = 0; i < (C < R ? C : R); i++) value[i][i] = s;
    }
    constexpr vec<R, T, Q>& operator[](int i) { return value[i]; }
    constexpr const vec<R, T, Q>& operator[](int i) const { return value[i]; }
};
using mat2 = mat<2, 2, float>;
using mat3 = mat<3, 3, float>;
using mat4 = mat<4, 4, float>;

template <typename T, qualifier Q>
constexpr vec<4, T, Q> operator*(const mat<4, 4, T, Q>& m, const vec<4, T, Q>& v) {
    return m[0] * v.x + m[1] * v.y + m[2] * v.z + m[3] * v.w;
}
template <typename T, qualifier Q>
constexpr mat<4, 4, T, Q> operator*(
    const mat<4, 4, T, Q>& a, const mat<4, 4, T, Q>& b
) {
    mat<4, 4, T, Q> r;
    for (int c = 0; c < 4; c++) r[c] = a * b[c];
    return r;
}
template <typename T, qualifier Q>
constexpr bool operator==(const mat<4, 4, T, Q>& a, const mat<4, 4, T, Q>& b) {
    return a[0] == b[0] && a[1] == b[1] && a[2] == b[2] && a[3] == b[3];
}
template <typename T, qualifier Q>
inline mat<4, 4, T, Q> translate(const mat<4, 4, T, Q>& m, const vec<3, T, Q>& v) {
    mat<4, 4, T, Q> r(m);
    r[3] = m[0] * v.x + m[1] * v.y + m[2] * v.z + m[3];
    return r;
}
template <typename T, qualifier Q>
inline mat<4, 4, T, Q> scale(const mat<4, 4, T, Q>& m, const vec<3, T, Q>& v) {
    mat<4, 4, T, Q> r;
    r[0] = m[0] * v.x;
    r[1] = m[1] * v.y;
    r[2] = m[2] * v.z;
    r[3] = m[3];
    return r;
}
template <typename T, qualifier Q>
inline mat<4, 4, T, Q> inverse(const mat<4, 4, T, Q>& m) {
    return m;  // never evaluated on the light/mesh path
}
template <typename T, qualifier Q>
inline mat<4, 4, T, Q> transpose(const mat<4, 4, T, Q>& m) {
    mat<4, 4, T, Q> r;
    for (int c = 0; c < 4; c++)
        for (int k = 0; k < 4; k++) r[c][k] = m[k][c];
    return r;
}

template <typename T, qualifier Q = defaultp>
struct qua {
    T x, y, z, w;
    constexpr qua() : x(0), y(0), z(0), w(1) {}
    constexpr qua(T w_, T x_, T y_, T z_) : x(x_), y(y_), z(z_), w(w_) {}
};
using quat = qua<float>;

template <typename T>
inline const T* value_ptr(const mat<4, 4, T>& m) {
    return &m[0].x;
}
template <int L, typename T, qualifier Q>
inline const T* value_ptr(const vec<L, T, Q>& v) {
    return &v.x;
}

}  // namespace glm

namespace std {
template <>
struct hash<glm::ivec2> {
    size_t operator()(const glm::ivec2& v) const noexcept {
        size_t seed = 0;
        hash<int> h;
        seed ^= h(v.x) + 0x9e3779b9 + (seed << 6) + (seed >> 2);
        seed ^= h(v.y) + 0x9e3779b9 + (seed << 6) + (seed >> 2);
        return seed;
    }
};
template <>
struct hash<glm::ivec3> {
    size_t operator()(const glm::ivec3& v) const noexcept {
        size_t seed = 0;
        hash<int> h;
        seed ^= h(v.x) + 0x9e3779b9 + (seed << 6) + (seed >> 2);
        seed ^= h(v.y) + 0x9e3779b9 + (seed << 6) + (seed >> 2);
        seed ^= h(v.z) + 0x9e3779b9 + (seed << 6) + (seed >> 2);
        return seed;
    }
};
}  // namespace std
