// gsmath.h -- the handful of glm 1.0.0 operations the reference's host side uses on the hot
// path (src/Renderer.cpp:46-49,80,726-736; src/GSScene.cpp:42-45), restated with glm's
// evaluation order so the uniform block is reproducible bit for bit.  glm is a FetchContent
// dependency of the reference (CMakeLists.txt:31-36) and is not vendored, so this is our own
// minimal implementation of its published formulas, not a copy.  Column-major like glm.
#pragma once
#include <cmath>
#include <cstring>

namespace gsmath {

struct vec3 {
    float x = 0, y = 0, z = 0;
};
struct vec4 {
    float x = 0, y = 0, z = 0, w = 0;
};
struct quat {  // glm::quat(w, x, y, z)
    float w = 1, x = 0, y = 0, z = 0;
};
struct mat4 {
    float m[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};  // m[col * 4 + row]
    float& at(int c, int r) { return m[c * 4 + r]; }
    float at(int c, int r) const { return m[c * 4 + r]; }
};

inline vec3 operator+(vec3 a, vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline vec3 operator*(vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline vec3 cross(vec3 a, vec3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
inline float radians(float deg) { return deg * 0.01745329251994329576923690768489f; }

// glm operator*(quat, vec3)
inline vec3 rotate(quat q, vec3 v) {
    const vec3 qv{q.x, q.y, q.z};
    const vec3 uv = cross(qv, v);
    const vec3 uuv = cross(qv, uv);
    return v + ((uv * q.w) + uuv) * 2.0f;
}
// glm operator*(quat, quat)
inline quat operator*(quat p, quat q) {
    quat r;
    r.w = p.w * q.w - p.x * q.x - p.y * q.y - p.z * q.z;
    r.x = p.w * q.x + p.x * q.w + p.y * q.z - p.z * q.y;
    r.y = p.w * q.y + p.y * q.w + p.z * q.x - p.x * q.z;
    r.z = p.w * q.z + p.z * q.w + p.x * q.y - p.y * q.x;
    return r;
}
// glm::rotate(quat, angle, axis) (gtc/quaternion)
inline quat rotate(quat q, float angle, vec3 axis) {
    vec3 t = axis;
    const float len = std::sqrt(t.x * t.x + t.y * t.y + t.z * t.z);
    if (std::fabs(len - 1.0f) > 0.001f) {
        const float inv = 1.0f / len;
        t = t * inv;
    }
    const float s = std::sin(angle * 0.5f);
    return q * quat{std::cos(angle * 0.5f), t.x * s, t.y * s, t.z * s};
}
// glm::normalize(vec4): v * inversesqrt(dot(v, v)), dot = (x*x + y*y) + (z*z + w*w)
inline vec4 normalize(vec4 v) {
    const float d = (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    const float inv = 1.0f / std::sqrt(d);
    return {v.x * inv, v.y * inv, v.z * inv, v.w * inv};
}

inline mat4 mat4_cast(quat q) {
    mat4 r;
    const float qxx = q.x * q.x, qyy = q.y * q.y, qzz = q.z * q.z;
    const float qxz = q.x * q.z, qxy = q.x * q.y, qyz = q.y * q.z;
    const float qwx = q.w * q.x, qwy = q.w * q.y, qwz = q.w * q.z;
    r.at(0, 0) = 1.0f - 2.0f * (qyy + qzz);
    r.at(0, 1) = 2.0f * (qxy + qwz);
    r.at(0, 2) = 2.0f * (qxz - qwy);
    r.at(1, 0) = 2.0f * (qxy - qwz);
    r.at(1, 1) = 1.0f - 2.0f * (qxx + qzz);
    r.at(1, 2) = 2.0f * (qyz + qwx);
    r.at(2, 0) = 2.0f * (qxz + qwy);
    r.at(2, 1) = 2.0f * (qyz - qwx);
    r.at(2, 2) = 1.0f - 2.0f * (qxx + qyy);
    return r;
}

inline mat4 translate(const mat4& m, vec3 v) {
    mat4 r = m;
    for (int row = 0; row < 4; row++) {
        float s = m.at(0, row) * v.x;
        s = s + m.at(1, row) * v.y;
        s = s + m.at(2, row) * v.z;
        r.at(3, row) = s + m.at(3, row);
    }
    return r;
}

inline mat4 operator*(const mat4& a, const mat4& b) {
    mat4 r;
    for (int c = 0; c < 4; c++)
        for (int row = 0; row < 4; row++) {
            float s = a.at(0, row) * b.at(c, 0);
            s = s + a.at(1, row) * b.at(c, 1);
            s = s + a.at(2, row) * b.at(c, 2);
            s = s + a.at(3, row) * b.at(c, 3);
            r.at(c, row) = s;
        }
    return r;
}

// glm::inverse(mat4): cofactor expansion in glm's operand order
inline mat4 inverse(const mat4& m) {
    auto M = [&](int c, int r) { return m.at(c, r); };
    const float s00 = M(2, 2) * M(3, 3) - M(3, 2) * M(2, 3), s02 = M(1, 2) * M(3, 3) - M(3, 2) * M(1, 3),
                s03 = M(1, 2) * M(2, 3) - M(2, 2) * M(1, 3), s04 = M(2, 1) * M(3, 3) - M(3, 1) * M(2, 3),
                s06 = M(1, 1) * M(3, 3) - M(3, 1) * M(1, 3), s07 = M(1, 1) * M(2, 3) - M(2, 1) * M(1, 3),
                s08 = M(2, 1) * M(3, 2) - M(3, 1) * M(2, 2), s10 = M(1, 1) * M(3, 2) - M(3, 1) * M(1, 2),
                s11 = M(1, 1) * M(2, 2) - M(2, 1) * M(1, 2), s12 = M(2, 0) * M(3, 3) - M(3, 0) * M(2, 3),
                s14 = M(1, 0) * M(3, 3) - M(3, 0) * M(1, 3), s15 = M(1, 0) * M(2, 3) - M(2, 0) * M(1, 3),
                s16 = M(2, 0) * M(3, 2) - M(3, 0) * M(2, 2), s18 = M(1, 0) * M(3, 2) - M(3, 0) * M(1, 2),
                s19 = M(1, 0) * M(2, 2) - M(2, 0) * M(1, 2), s20 = M(2, 0) * M(3, 1) - M(3, 0) * M(2, 1),
                s22 = M(1, 0) * M(3, 1) - M(3, 0) * M(1, 1), s23 = M(1, 0) * M(2, 1) - M(2, 0) * M(1, 1);
    const float F0[4] = {s00, s00, s02, s03}, F1[4] = {s04, s04, s06, s07}, F2[4] = {s08, s08, s10, s11};
    const float F3[4] = {s12, s12, s14, s15}, F4[4] = {s16, s16, s18, s19}, F5[4] = {s20, s20, s22, s23};
    const float V0[4] = {M(1, 0), M(0, 0), M(0, 0), M(0, 0)}, V1[4] = {M(1, 1), M(0, 1), M(0, 1), M(0, 1)};
    const float V2[4] = {M(1, 2), M(0, 2), M(0, 2), M(0, 2)}, V3[4] = {M(1, 3), M(0, 3), M(0, 3), M(0, 3)};
    mat4 inv;
    for (int i = 0; i < 4; i++) {
        const float sgnA = (i & 1) ? -1.0f : 1.0f, sgnB = -sgnA;
        inv.at(0, i) = ((V1[i] * F0[i] - V2[i] * F1[i]) + V3[i] * F2[i]) * sgnA;
        inv.at(1, i) = ((V0[i] * F0[i] - V2[i] * F3[i]) + V3[i] * F4[i]) * sgnB;
        inv.at(2, i) = ((V0[i] * F1[i] - V1[i] * F3[i]) + V3[i] * F5[i]) * sgnA;
        inv.at(3, i) = ((V0[i] * F2[i] - V1[i] * F4[i]) + V2[i] * F5[i]) * sgnB;
    }
    const float d0 = M(0, 0) * inv.at(0, 0), d1 = M(0, 1) * inv.at(1, 0), d2 = M(0, 2) * inv.at(2, 0),
                d3 = M(0, 3) * inv.at(3, 0);
    const float ood = 1.0f / ((d0 + d1) + (d2 + d3));
    for (float& f : inv.m) f = f * ood;
    return inv;
}

// glm::perspective with default defines == perspectiveRH_NO
inline mat4 perspective(float fovy, float aspect, float z_near, float z_far) {
    const float tan_half = std::tan(fovy / 2.0f);
    mat4 r;
    std::memset(r.m, 0, sizeof r.m);
    r.at(0, 0) = 1.0f / (aspect * tan_half);
    r.at(1, 1) = 1.0f / tan_half;
    r.at(2, 2) = -(z_far + z_near) / (z_far - z_near);
    r.at(2, 3) = -1.0f;
    r.at(3, 2) = -(2.0f * z_far * z_near) / (z_far - z_near);
    return r;
}

}  // namespace gsmath
