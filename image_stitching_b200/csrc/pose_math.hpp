// pose_math.hpp - host-side camera-rotation helpers of the compositing contract.
//
// Restates, on plain row-major 3x3 arrays (no cv::Mat), the members of the reference's
// Quaternion<T> (image_stitching/quaternion.h) and the euler conversions
// (image_stitching/euler.h, euler_order.h) that feed K/R into the warper.  The arithmetic
// (operation order) is kept identical so results are bit-equal with the reference's templates
// instantiated on the same float type; the code organisation is new.
#pragma once
#include <algorithm>
#include <cmath>
#include <limits>

namespace isb {

enum class EulerOrder { XYZ = 0, YXZ, ZXY, ZYX, YZX, XZY };  // euler_order.h:3-11

// m is row-major: m[3*r + c]
template <typename T>
struct Quat {
    T x{0}, y{0}, z{0}, w{1};

    // quaternion.h:260-322 (setFromRotationMatrix)
    static Quat from_rotation(const T* m)
    {
        const T m11 = m[0], m12 = m[1], m13 = m[2], m21 = m[3], m22 = m[4], m23 = m[5], m31 = m[6], m32 = m[7], m33 = m[8];
        Quat q;
        const T trace = m11 + m22 + m33;
        if (trace > 0) {
            const auto s = 0.5 / std::sqrt(trace + 1.0);
            q.w = T(0.25 / s);
            q.x = T((m32 - m23) * s);
            q.y = T((m13 - m31) * s);
            q.z = T((m21 - m12) * s);
        } else if (m11 > m22 && m11 > m33) {
            const auto s = 2.0 * std::sqrt(1.0 + m11 - m22 - m33);
            q.w = T((m32 - m23) / s);
            q.x = T(0.25 * s);
            q.y = T((m12 + m21) / s);
            q.z = T((m13 + m31) / s);
        } else if (m22 > m33) {
            const auto s = 2.0 * std::sqrt(1.0 + m22 - m11 - m33);
            q.w = T((m13 - m31) / s);
            q.x = T((m12 + m21) / s);
            q.y = T(0.25 * s);
            q.z = T((m23 + m32) / s);
        } else {
            const auto s = 2.0 * std::sqrt(1.0 + m33 - m11 - m22);
            q.w = T((m21 - m12) / s);
            q.x = T((m13 + m31) / s);
            q.y = T((m23 + m32) / s);
            q.z = T(0.25 * s);
        }
        return q;
    }

    // quaternion.h:564-596 (toRotationMatrix)
    void to_rotation(T* m) const
    {
        const T x2 = x + x, y2 = y + y, z2 = z + z;
        const T xx = x * x2, xy = x * y2, xz = x * z2;
        const T yy = y * y2, yz = y * z2, zz = z * z2;
        const T wx = w * x2, wy = w * y2, wz = w * z2;
        m[0] = (1 - (yy + zz)); m[1] = (xy - wz);       m[2] = (xz + wy);
        m[3] = (xy + wz);       m[4] = (1 - (xx + zz)); m[5] = (yz - wx);
        m[6] = (xz - wy);       m[7] = (yz + wx);       m[8] = (1 - (xx + yy));
    }

    // quaternion.h:172-239 (setFromEuler)
    static Quat from_euler(T ex, T ey, T ez, EulerOrder order)
    {
        const auto c1 = std::cos(ex / 2), c2 = std::cos(ey / 2), c3 = std::cos(ez / 2);
        const auto s1 = std::sin(ex / 2), s2 = std::sin(ey / 2), s3 = std::sin(ez / 2);
        // sign pattern of the four cross terms per order: {x, y, z, w}
        static const int sg[6][4] = {{+1, -1, +1, -1},   // XYZ
                                     {+1, -1, -1, +1},   // YXZ
                                     {-1, +1, +1, -1},   // ZXY
                                     {-1, +1, -1, +1},   // ZYX
                                     {+1, +1, -1, -1},   // YZX
                                     {-1, -1, +1, +1}};  // XZY
        const int* s = sg[int(order)];
        Quat q;
        q.x = s[0] > 0 ? s1 * c2 * c3 + c1 * s2 * s3 : s1 * c2 * c3 - c1 * s2 * s3;
        q.y = s[1] > 0 ? c1 * s2 * c3 + s1 * c2 * s3 : c1 * s2 * c3 - s1 * c2 * s3;
        q.z = s[2] > 0 ? c1 * c2 * s3 + s1 * s2 * c3 : c1 * c2 * s3 - s1 * s2 * c3;
        q.w = s[3] > 0 ? c1 * c2 * c3 + s1 * s2 * s3 : c1 * c2 * c3 - s1 * s2 * s3;
        return q;
    }

    // quaternion.h:241-258 (setFromAxisAngle; axis assumed normalised)
    static Quat from_axis_angle(const T* axis, T angle)
    {
        const auto half = angle / 2;
        const auto s = std::sin(half);
        Quat q;
        q.x = axis[0] * s; q.y = axis[1] * s; q.z = axis[2] * s; q.w = std::cos(half);
        return q;
    }

    // quaternion.h:464-478 (multiplyQuaternions)
    static Quat multiply(const Quat& a, const Quat& b)
    {
        Quat q;
        q.x = a.x * b.w + a.w * b.x + a.y * b.z - a.z * b.y;
        q.y = a.y * b.w + a.w * b.y + a.z * b.x - a.x * b.z;
        q.z = a.z * b.w + a.w * b.z + a.x * b.y - a.y * b.x;
        q.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
        return q;
    }

    T length() const { return std::sqrt(x * x + y * y + z * z + w * w); }

    void normalize()
    {
        T l = length();
        if (l == 0) { x = y = z = 0; w = 1; return; }
        l = T(1) / l;
        x = x * l; y = y * l; z = z * l; w = w * l;
    }

    // quaternion.h:480-544 (slerp of *this towards qb)
    Quat slerp(const Quat& qb, T t) const
    {
        if (t == 0) return *this;
        if (t == 1) return qb;
        Quat r;
        auto cos_half = w * qb.w + x * qb.x + y * qb.y + z * qb.z;
        if (cos_half < 0) { r.w = -qb.w; r.x = -qb.x; r.y = -qb.y; r.z = -qb.z; cos_half = -cos_half; }
        else r = qb;
        if (cos_half >= 1.0) return *this;
        const auto sqr_sin = 1.0 - cos_half * cos_half;
        if (sqr_sin <= std::numeric_limits<T>::epsilon()) {
            const auto s = 1 - t;
            r.w = s * w + t * r.w; r.x = s * x + t * r.x; r.y = s * y + t * r.y; r.z = s * z + t * r.z;
            r.normalize();
            return r;
        }
        const auto sin_half = std::sqrt(sqr_sin);
        const auto half = std::atan2(sin_half, cos_half);
        const auto ra = std::sin((1 - t) * half) / sin_half, rb = std::sin(t * half) / sin_half;
        r.w = T(w * ra + r.w * rb); r.x = T(x * ra + r.x * rb); r.y = T(y * ra + r.y * rb); r.z = T(z * ra + r.z * rb);
        return r;
    }
};

// euler.h:4-133.  Returns false for an unknown order.
template <typename T>
bool rotation_to_euler(const T* m, EulerOrder order, T* e)
{
    const T m11 = m[0], m12 = m[1], m13 = m[2], m21 = m[3], m22 = m[4], m23 = m[5], m31 = m[6], m32 = m[7], m33 = m[8];
    auto clampT = [](T v) { return std::clamp(v, T(-1), T(1)); };
    const double lim = 0.9999999;
    T x, y, z;
    switch (order) {
    case EulerOrder::XYZ:
        y = std::asin(clampT(m13));
        if (std::abs(m13) < lim) { x = std::atan2(-m23, m33); z = std::atan2(-m12, m11); }
        else { x = std::atan2(m32, m22); z = 0; }
        break;
    case EulerOrder::YXZ:
        x = std::asin(-clampT(m23));
        if (std::abs(m23) < lim) { y = std::atan2(m13, m33); z = std::atan2(m21, m22); }
        else { y = std::atan2(-m31, m11); z = 0; }
        break;
    case EulerOrder::ZXY:
        x = std::asin(clampT(m32));
        if (std::abs(m32) < lim) { y = std::atan2(-m31, m33); z = std::atan2(-m12, m22); }
        else { y = 0; z = std::atan2(m21, m11); }
        break;
    case EulerOrder::ZYX:
        y = std::asin(-clampT(m31));
        if (std::abs(m31) < lim) { x = std::atan2(m32, m33); z = std::atan2(m21, m11); }
        else { x = 0; z = std::atan2(-m12, m22); }
        break;
    case EulerOrder::YZX:
        z = std::asin(clampT(m21));
        if (std::abs(m21) < lim) { x = std::atan2(-m23, m22); y = std::atan2(-m31, m11); }
        else { x = 0; y = std::atan2(m13, m33); }
        break;
    case EulerOrder::XZY:
        z = std::asin(-clampT(m12));
        if (std::abs(m12) < lim) { x = std::atan2(m32, m22); y = std::atan2(m13, m11); }
        else { x = std::atan2(-m23, m33); y = 0; }
        break;
    default:
        return false;
    }
    e[0] = x; e[1] = y; e[2] = z;
    return true;
}

// euler.h:135-300 (three.js makeRotationFromEuler; the reference stores te[0],te[4],te[8] across row 0, :289-297).
template <typename T>
bool euler_to_rotation(const T* eul, EulerOrder order, T* m)
{
    const T a = std::cos(eul[0]), b = std::sin(eul[0]);
    const T c = std::cos(eul[1]), d = std::sin(eul[1]);
    const T e = std::cos(eul[2]), f = std::sin(eul[2]);
    switch (order) {
    case EulerOrder::XYZ: {
        const T ae = a * e, af = a * f, be = b * e, bf = b * f;
        m[0] = c * e;        m[1] = -c * f;       m[2] = d;
        m[3] = af + be * d;  m[4] = ae - bf * d;  m[5] = -b * c;
        m[6] = bf - ae * d;  m[7] = be + af * d;  m[8] = a * c;
        break; }
    case EulerOrder::YXZ: {
        const T ce = c * e, cf = c * f, de = d * e, df = d * f;
        m[0] = ce + df * b;  m[1] = de * b - cf;  m[2] = a * d;
        m[3] = a * f;        m[4] = a * e;        m[5] = -b;
        m[6] = cf * b - de;  m[7] = df + ce * b;  m[8] = a * c;
        break; }
    case EulerOrder::ZXY: {
        const T ce = c * e, cf = c * f, de = d * e, df = d * f;
        m[0] = ce - df * b;  m[1] = -a * f;       m[2] = de + cf * b;
        m[3] = cf + de * b;  m[4] = a * e;        m[5] = df - ce * b;
        m[6] = -a * d;       m[7] = b;            m[8] = a * c;
        break; }
    case EulerOrder::ZYX: {
        const T ae = a * e, af = a * f, be = b * e, bf = b * f;
        m[0] = c * e;        m[1] = be * d - af;  m[2] = ae * d + bf;
        m[3] = c * f;        m[4] = bf * d + ae;  m[5] = af * d - be;
        m[6] = -d;           m[7] = b * c;        m[8] = a * c;
        break; }
    case EulerOrder::YZX: {
        const T ac = a * c, ad = a * d, bc = b * c, bd = b * d;
        m[0] = c * e;        m[1] = bd - ac * f;  m[2] = bc * f + ad;
        m[3] = f;            m[4] = a * e;        m[5] = -b * e;
        m[6] = -d * e;       m[7] = ad * f + bc;  m[8] = ac - bd * f;
        break; }
    case EulerOrder::XZY: {
        const T ac = a * c, ad = a * d, bc = b * c, bd = b * d;
        m[0] = c * e;        m[1] = -f;           m[2] = d * e;
        m[3] = ac * f + bd;  m[4] = a * e;        m[5] = ad * f - bc;
        m[6] = bc * f - ad;  m[7] = b * e;        m[8] = bd * f + ac;
        break; }
    default:
        return false;
    }
    return true;
}

}  // namespace isb
