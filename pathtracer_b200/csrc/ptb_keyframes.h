// ptb_keyframes.h — key-framed object placement (host code): Object::get_translation / get_rotation / get_scale
// (Geometry.h:258-312) with the reference's quaternion Slerp between rotation keys (Vector.h:104-160, 222-269).
#pragma once
#include <math.h>

#include <algorithm>
#include <vector>

namespace ptb {

struct KeyTrack {              // one std::map<float, value> of the reference: frames ascending, one value row per frame
    std::vector<float> frames;
    std::vector<float> values; // frames.size() x width
    int width = 1;             // 1 scale, 3 translation, 9 rotation (row-major Matrix33)
    bool empty() const { return frames.empty(); }
};

// std::map semantics: ascending keys, a later assignment to the same frame replaces the earlier one
inline void key_track_set(KeyTrack& k, const float* frames, const float* values, int n, int width) {
    k.width = width; k.frames.clear(); k.values.clear();
    std::vector<int> order(n);
    for (int i = 0; i < n; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return frames[a] < frames[b]; });
    for (int i = 0; i < n; i++) {
        const int s = order[i];
        if (!k.frames.empty() && k.frames.back() == frames[s]) { std::copy(values + (size_t)s * width, values + (size_t)(s + 1) * width, k.values.end() - width); continue; }
        k.frames.push_back(frames[s]);
        k.values.insert(k.values.end(), values + (size_t)s * width, values + (size_t)(s + 1) * width);
    }
}

// Matrix<3,3,float>::toQuaternion (Vector.h:119-160): note m01 = (*this)(1,0) etc.; `sqrt(tr + 1.0) * 2` is evaluated in double
inline void mat_to_quat(const float* v, float q[4]) {
    const float m00 = v[0], m01 = v[3], m02 = v[6], m10 = v[1], m11 = v[4], m12 = v[7], m20 = v[2], m21 = v[5], m22 = v[8];
    const float tr = m00 + m11 + m22;
    float qw, qx, qy, qz;
    if (tr > 0) {
        const float S = (float)(sqrt(tr + 1.0) * 2);
        qw = (float)(0.25 * S); qx = (m21 - m12) / S; qy = (m02 - m20) / S; qz = (m10 - m01) / S;
    } else if ((m00 > m11) & (m00 > m22)) {
        const float S = (float)(sqrt(1.0 + m00 - m11 - m22) * 2);
        qw = (m21 - m12) / S; qx = (float)(0.25 * S); qy = (m01 + m10) / S; qz = (m02 + m20) / S;
    } else if (m11 > m22) {
        const float S = (float)(sqrt(1.0 + m11 - m00 - m22) * 2);
        qw = (m02 - m20) / S; qx = (m01 + m10) / S; qy = (float)(0.25 * S); qz = (m12 + m21) / S;
    } else {
        const float S = (float)(sqrt(1.0 + m22 - m00 - m11) * 2);
        qw = (m10 - m01) / S; qx = (m02 + m20) / S; qy = (m12 + m21) / S; qz = (float)(0.25 * S);
    }
    q[0] = qw; q[1] = qx; q[2] = qy; q[3] = qz;
}
// Matrix::fromQuaternion (Vector.h:104-117): the `2.0*x*y` products are double
inline void quat_to_mat(const float q[4], float* v) {
    const float w = q[0], x = q[1], y = q[2], z = q[3];
    v[0] = w * w + x * x - y * y - z * z;
    v[1] = (float)(2.0 * x * y + 2.0 * w * z);
    v[2] = (float)(2.0 * x * z - 2.0 * y * w);
    v[3] = (float)(2.0 * x * y - 2.0 * w * z);
    v[4] = w * w - x * x + y * y - z * z;
    v[5] = (float)(2.0 * y * z + 2.0 * w * x);
    v[6] = (float)(2.0 * x * z + 2.0 * w * y);
    v[7] = (float)(2.0 * y * z - 2.0 * w * x);
    v[8] = w * w - x * x - y * y + z * z;
}
// Slerp(Matrix33, Matrix33, t) (Vector.h:222-269)
inline void slerp33(const float* a, const float* b, float t, float* out) {
    float q1[4], q2[4];
    mat_to_quat(a, q1); mat_to_quat(b, q2);
    float w2 = q2[0], x2 = q2[1], y2 = q2[2], z2 = q2[3];
    const float w1 = q1[0], x1 = q1[1], y1 = q1[2], z1 = q1[3];
    if (w1 * w2 + x1 * x2 + y1 * y2 + z1 * z2 < 0) { w2 = -w2; x2 = -x2; y2 = -y2; z2 = -z2; }
    const float theta = acosf(w1 * w2 + x1 * x2 + y1 * y2 + z1 * z2);
    float mult1, mult2;
    if (theta > 0.000001) { mult1 = sinf((1 - t) * theta) / sinf(theta); mult2 = sinf(t * theta) / sinf(theta); }
    else { mult1 = 1 - t; mult2 = t; }
    const float q3[4] = {mult1 * w1 + mult2 * w2, mult1 * x1 + mult2 * x2, mult1 * y1 + mult2 * y2, mult1 * z1 + mult2 * z2};
    quat_to_mat(q3, out);
}
// get_translation / get_scale / get_rotation at `frame`; false when the track has no key (the static placement applies)
inline bool key_eval(const KeyTrack& k, float frame, float* out) {
    const size_t n = k.frames.size();
    if (n == 0) return false;
    const int w = k.width;
    size_t up = 0;
    while (up < n && !(k.frames[up] > frame)) up++;                    // upper_bound(frame)
    if (up == n) { std::copy(k.values.end() - w, k.values.end(), out); return true; }
    if (up == 0) { std::copy(k.values.begin(), k.values.begin() + w, out); return true; }
    const float t = (frame - k.frames[up - 1]) / (k.frames[up] - k.frames[up - 1]);
    const float* a = &k.values[(up - 1) * w];
    const float* b = &k.values[up * w];
    if (w == 9) slerp33(a, b, t, out);
    else if (w == 3) for (int i = 0; i < 3; i++) out[i] = (1 - t) * a[i] + t * b[i];
    else out[0] = (1.f - t) * a[0] + t * b[0];
    return true;
}

}  // namespace ptb
