// Pupil::util — host-side maths and plumbing of the kept surface, re-stated without DirectXMath / spdlog.
//
//   Float3 / Float4 / Mat4      framework/util/type.h:7-112     (row-major storage, column vectors: p' = M p)
//   Transform                   framework/util/transform.{h,cpp} (Scale / Rotate / Translate PRE-multiply)
//   Camera / CameraDesc         framework/util/camera.{h,cpp}
//   AABB                        framework/util/aabb.h
//   Singleton, Timer, Log       framework/util/{util.h,timer.h,log.h}
//   Event                       framework/util/event.h  (EventBinder<e>(fn) / EventDispatcher<e>(payload))
//
// DirectXMath (Windows SDK) is what the reference calls for perspective / look-at / inverse; its
// published definitions are written out here in plain fp32, inverses in fp64 rounded once.
#pragma once
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <functional>
#include <mutex>
#include <string>
#include <string_view>
#include <tuple>
#include <type_traits>
#include <unordered_map>
#include <vector>

namespace Pupil {
namespace util {

template<typename T>
class Singleton {
public:
    static T *instance() {
        static T inst;
        return &inst;
    }
    Singleton() = default;
    Singleton(const Singleton &) = delete;
    Singleton &operator=(const Singleton &) = delete;
};

struct Float3 {
    union {
        struct {
            float x, y, z;
        };
        struct {
            float r, g, b;
        };
        float e[3];
    };
    constexpr Float3(float x_, float y_, float z_) noexcept : x(x_), y(y_), z(z_) {}
    constexpr Float3(float v = 0.f) noexcept : x(v), y(v), z(v) {}
};
struct Float4 {
    float x = 0.f, y = 0.f, z = 0.f, w = 0.f;
};

struct Mat4 {
    union {
        float e[16];
        float re[4][4];
    };
    Mat4() noexcept : Mat4(1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1) {}
    Mat4(float a0, float a1, float a2, float a3, float b0, float b1, float b2, float b3, float c0, float c1, float c2, float c3, float d0,
         float d1, float d2, float d3) noexcept
        : e{ a0, a1, a2, a3, b0, b1, b2, b3, c0, c1, c2, c3, d0, d1, d2, d3 } {}
    Mat4 GetTranspose() const noexcept {
        Mat4 t;
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) t.re[i][j] = re[j][i];
        return t;
    }
    // general inverse: Gauss-Jordan with partial pivoting in fp64, rounded once to fp32
    Mat4 GetInverse() const noexcept {
        double a[4][8];
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) a[i][j] = re[i][j], a[i][j + 4] = (i == j);
        for (int col = 0; col < 4; ++col) {
            int piv = col;
            for (int row = col + 1; row < 4; ++row)
                if (std::fabs(a[row][col]) > std::fabs(a[piv][col])) piv = row;
            if (piv != col)
                for (int k = 0; k < 8; ++k) std::swap(a[piv][k], a[col][k]);
            const double d = a[col][col];
            if (d == 0.0) continue;
            for (int k = 0; k < 8; ++k) a[col][k] /= d;
            for (int row = 0; row < 4; ++row) {
                if (row == col || a[row][col] == 0.0) continue;
                const double f = a[row][col];
                for (int k = 0; k < 8; ++k) a[row][k] -= f * a[col][k];
            }
        }
        Mat4 out;
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) out.re[i][j] = static_cast<float>(a[i][j + 4]);
        return out;
    }
};
inline Mat4 operator*(const Mat4 &a, const Mat4 &b) noexcept {
    Mat4 out;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            float s = 0.f;
            for (int k = 0; k < 4; ++k) s += a.re[i][k] * b.re[k][j];
            out.re[i][j] = s;
        }
    return out;
}

struct Transform {
    Mat4 matrix;
    Transform() noexcept = default;
    Transform(const Mat4 &m) noexcept : matrix(m) {}

    void Translate(float x, float y, float z) noexcept { // transform.cpp:7-15
        Mat4 t;
        t.re[0][3] = x, t.re[1][3] = y, t.re[2][3] = z;
        matrix = t * matrix;
    }
    void Rotate(float ux, float uy, float uz, float angle) noexcept { // transform.cpp:17-70, unit quaternion -> 3x3
        const float len = std::sqrt(ux * ux + uy * uy + uz * uz);
        ux /= len, uy /= len, uz /= len;
        const float theta = angle / 180.f * 3.14159265358979323846f;
        const float qw = std::cos(0.5f * theta);
        const float qx = std::sin(0.5f * theta) * ux, qy = std::sin(0.5f * theta) * uy, qz = std::sin(0.5f * theta) * uz;
        Mat4 rot;
        rot.re[0][0] = 1.f - 2.f * qy * qy - 2.f * qz * qz, rot.re[0][1] = 2.f * qx * qy - 2.f * qw * qz, rot.re[0][2] = 2.f * qw * qy + 2.f * qx * qz;
        rot.re[1][0] = 2.f * qx * qy + 2.f * qw * qz, rot.re[1][1] = 1.f - 2.f * qx * qx - 2.f * qz * qz, rot.re[1][2] = 2.f * qy * qz - 2.f * qw * qx;
        rot.re[2][0] = 2.f * qx * qz - 2.f * qw * qy, rot.re[2][1] = 2.f * qw * qx + 2.f * qy * qz, rot.re[2][2] = 1.f - 2.f * qx * qx - 2.f * qy * qy;
        matrix = rot * matrix;
    }
    void Scale(float x, float y, float z) noexcept { // transform.cpp:72-84
        Mat4 s;
        s.re[0][0] = x, s.re[1][1] = y, s.re[2][2] = z;
        matrix = s * matrix;
    }
    // transform.cpp:86-97: matrix = transpose(inverse(XMMatrixLookAtRH(origin, target, up))) in DirectX's
    // row-vector storage, i.e. the camera-to-world matrix for column vectors.
    // XMMatrixLookAtRH: z = normalize(origin - target), x = normalize(cross(up, z)), y = cross(z, x).
    void LookAt(const Float3 &origin, const Float3 &target, const Float3 &up) noexcept {
        auto norm = [](Float3 v) {
            const float l = 1.f / std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
            return Float3{ v.x * l, v.y * l, v.z * l };
        };
        auto crs = [](Float3 a, Float3 b) { return Float3{ a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x }; };
        auto dt = [](Float3 a, Float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; };
        const Float3 zc = norm(Float3{ origin.x - target.x, origin.y - target.y, origin.z - target.z });
        const Float3 xc = norm(crs(up, zc));
        const Float3 yc = crs(zc, xc);
        Mat4 view(xc.x, xc.y, xc.z, -dt(xc, origin), yc.x, yc.y, yc.z, -dt(yc, origin), zc.x, zc.y, zc.z, -dt(zc, origin), 0, 0, 0, 1);
        matrix = view.GetInverse();
    }
    static Float3 TransformPoint(const Float3 p, const Mat4 &m) noexcept { // transform.cpp:99-107
        const float x = m.e[0] * p.x + m.e[1] * p.y + m.e[2] * p.z + m.e[3];
        const float y = m.e[4] * p.x + m.e[5] * p.y + m.e[6] * p.z + m.e[7];
        const float z = m.e[8] * p.x + m.e[9] * p.y + m.e[10] * p.z + m.e[11];
        const float w = m.e[12] * p.x + m.e[13] * p.y + m.e[14] * p.z + m.e[15];
        return Float3{ x / w, y / w, z / w };
    }
    static Float3 TransformVector(const Float3 v, const Mat4 &m) noexcept { // :109-114
        return Float3{ m.e[0] * v.x + m.e[1] * v.y + m.e[2] * v.z, m.e[4] * v.x + m.e[5] * v.y + m.e[6] * v.z,
                       m.e[8] * v.x + m.e[9] * v.y + m.e[10] * v.z };
    }
    static Float3 TransformNormal(const Float3 n, const Mat4 &inv_t) noexcept { // :116-123, normalised
        const Float3 v = TransformVector(n, inv_t);
        const float len = std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
        return Float3{ v.x / len, v.y / len, v.z / len };
    }
};

struct AABB {
    Float3 min{ 1e30f }, max{ -1e30f };
    void Merge(const Float3 &p) noexcept {
        min = Float3{ std::fmin(min.x, p.x), std::fmin(min.y, p.y), std::fmin(min.z, p.z) };
        max = Float3{ std::fmax(max.x, p.x), std::fmax(max.y, p.y), std::fmax(max.z, p.z) };
    }
    void Merge(const AABB &o) noexcept { Merge(o.min), Merge(o.max); }
    // box of the 8 transformed corners (util/aabb.h)
    void Transform(const Mat4 &m) noexcept {
        AABB out;
        for (int k = 0; k < 8; ++k)
            out.Merge(util::Transform::TransformPoint(Float3{ k & 1 ? max.x : min.x, k & 2 ? max.y : min.y, k & 4 ? max.z : min.z }, m));
        *this = out;
    }
};

struct CameraDesc {
    float fov_y = 45.f;
    float aspect_ratio = 1.f;
    float near_clip = 0.01f;
    float far_clip = 10000.f;
    Transform to_world;
};

class Camera {
    float m_fov_y = 45.f, m_aspect_ratio = 1.f, m_near_clip = 0.01f, m_far_clip = 10000.f;
    Float3 m_position;
    Mat4 m_rotate, m_rotate_inv;
    bool m_to_world_dirty = true;
    Mat4 m_to_world, m_view;
    bool m_projection_dirty = true;
    Mat4 m_sample_to_camera, m_proj;

    void UpdateProjection() noexcept {
        // XMMatrixPerspectiveFovRH(fov, aspect, zn, zf) for row vectors:
        //   [ h/aspect 0 0 0 ; 0 h 0 0 ; 0 0 zf/(zn-zf) -1 ; 0 0 zn*zf/(zn-zf) 0 ],  h = cot(fov/2)
        const float half = 0.5f * (m_fov_y / 180.f * 3.14159265358979323846f);
        const float h = std::cos(half) / std::sin(half);
        const float range = m_far_clip / (m_near_clip - m_far_clip);
        Mat4 proj(h / m_aspect_ratio, 0, 0, 0, 0, h, 0, 0, 0, 0, range, -1.f, 0, 0, range * m_near_clip, 0);
        Mat4 shift; // XMMatrixTranslation(1,1,0), row-vector form: translation in the last row
        shift.re[3][0] = 1.f, shift.re[3][1] = 1.f;
        Mat4 half_scale; // XMMatrixScaling(.5,.5,1)
        half_scale.re[0][0] = 0.5f, half_scale.re[1][1] = 0.5f;
        m_proj = proj.GetTranspose();
        m_sample_to_camera = ((proj * shift) * half_scale).GetInverse().GetTranspose(); // camera.cpp:9-16
        m_projection_dirty = false;
    }
    void UpdateToWorld() noexcept { // camera.cpp:35-47
        Mat4 t(1, 0, 0, -m_position.x, 0, 1, 0, -m_position.y, 0, 0, 1, -m_position.z, 0, 0, 0, 1);
        m_view = m_rotate * t;
        m_to_world = m_view.GetInverse();
        m_to_world_dirty = false;
    }

public:
    static inline float sensitivity = 0.05f;
    static inline float sensitivity_scale = 1.f;

    Float3 GetPosition() const noexcept { return m_position; }
    Mat4 GetSampleToCameraMatrix() noexcept {
        if (m_projection_dirty) UpdateProjection();
        return m_sample_to_camera;
    }
    Mat4 GetProjectionMatrix() noexcept {
        if (m_projection_dirty) UpdateProjection();
        return m_proj;
    }
    Mat4 GetToWorldMatrix() noexcept {
        if (m_to_world_dirty) UpdateToWorld();
        return m_to_world;
    }
    Mat4 GetViewMatrix() noexcept {
        if (m_to_world_dirty) UpdateToWorld();
        return m_view;
    }
    std::tuple<Float3, Float3, Float3> GetCameraCoordinateSystem() const noexcept {
        return { Transform::TransformVector(Float3{ 1, 0, 0 }, m_rotate_inv), Transform::TransformVector(Float3{ 0, 1, 0 }, m_rotate_inv),
                 Transform::TransformVector(Float3{ 0, 0, 1 }, m_rotate_inv) };
    }
    void SetProjectionFactor(float fov_y, float aspect_ratio, float near_clip = 0.01f, float far_clip = 10000.f) noexcept {
        m_fov_y = fov_y, m_aspect_ratio = aspect_ratio, m_near_clip = near_clip, m_far_clip = far_clip;
        m_projection_dirty = true;
    }
    void SetFov(float fov) noexcept { m_fov_y = fov, m_projection_dirty = true; }
    void SetWorldTransform(Transform to_world) noexcept { // camera.cpp:80-101: the matrix is kept verbatim
        m_to_world = to_world.matrix;
        m_position = Float3{ m_to_world.re[0][3], m_to_world.re[1][3], m_to_world.re[2][3] };
        m_rotate = to_world.matrix.GetTranspose();
        m_rotate.re[3][0] = m_rotate.re[3][1] = m_rotate.re[3][2] = 0.f;
        m_rotate_inv = m_rotate.GetTranspose();
        Mat4 t(1, 0, 0, -m_position.x, 0, 1, 0, -m_position.y, 0, 0, 1, -m_position.z, 0, 0, 0, 1);
        m_view = m_rotate * t;
        m_to_world_dirty = false;
    }
    void Rotate(float delta_x, float delta_y) noexcept { // camera.cpp:103-112
        Transform pitch, yaw;
        pitch.Rotate(1, 0, 0, delta_y);
        yaw.Rotate(0, 1, 0, delta_x);
        m_rotate = pitch.matrix * m_rotate * yaw.matrix;
        m_rotate_inv = m_rotate.GetTranspose();
        m_to_world_dirty = true;
    }
    void Move(Float3 delta) noexcept { // camera.cpp:114-120
        delta = Transform::TransformVector(delta, m_rotate_inv);
        m_position = Float3{ m_position.x + delta.x, m_position.y + delta.y, m_position.z + delta.z };
        m_to_world_dirty = true;
    }
};

inline std::vector<std::string> Split(std::string_view str, std::string_view delims) { // util/util.h Split: empty pieces are dropped
    std::vector<std::string> out;
    size_t i = 0;
    while (i < str.size()) {
        const size_t j = str.find_first_of(delims, i);
        const size_t end = j == std::string_view::npos ? str.size() : j;
        if (end > i) out.emplace_back(str.substr(i, end - i));
        i = end + 1;
    }
    return out;
}
}// namespace util

class Timer {
    std::chrono::steady_clock::time_point m_start{}, m_stop{};

public:
    void Start() noexcept { m_start = std::chrono::steady_clock::now(); }
    void Stop() noexcept { m_stop = std::chrono::steady_clock::now(); }
    double ElapsedMilliseconds() const noexcept { return std::chrono::duration<double, std::milli>(m_stop - m_start).count(); }
    double ElapsedSeconds() const noexcept { return ElapsedMilliseconds() * 1e-3; }
};

// printf-style logger (the reference formats with spdlog; messages are not part of the contract)
struct Log {
    static inline int level = 1; // 0 silent, 1 warn+error, 2 info
    static void Write(const char *tag, const char *fmt, va_list ap) {
        std::fprintf(stderr, "[pupil %s] ", tag);
        std::vfprintf(stderr, fmt, ap);
        std::fputc('\n', stderr);
    }
    static void Info(const char *fmt, ...) {
        if (level < 2) return;
        va_list ap;
        va_start(ap, fmt), Write("info", fmt, ap), va_end(ap);
    }
    static void Warn(const char *fmt, ...) {
        if (level < 1) return;
        va_list ap;
        va_start(ap, fmt), Write("warn", fmt, ap), va_end(ap);
    }
    static void Error(const char *fmt, ...) {
        if (level < 1) return;
        va_list ap;
        va_start(ap, fmt), Write("error", fmt, ap), va_end(ap);
    }
};

// ---- events: one handler list per enum VALUE (framework/util/event.h) -----------------------------------
template<auto E>
struct Event {
    using Handler = std::function<void(void *)>;
    static std::vector<Handler> &Handlers() {
        static std::vector<Handler> h;
        return h;
    }
    static std::mutex &Mutex() {
        static std::mutex m;
        return m;
    }
};
template<auto E>
inline void EventBinder(std::function<void(void *)> fn) {
    std::lock_guard lock(Event<E>::Mutex());
    Event<E>::Handlers().push_back(std::move(fn));
}
template<auto E>
inline void EventDispatcher(void *payload = nullptr) {
    std::vector<std::function<void(void *)>> copy;
    {
        std::lock_guard lock(Event<E>::Mutex());
        copy = Event<E>::Handlers();
    }
    for (auto &h : copy) h(payload);
}
template<auto E, typename T>
    requires(!std::is_pointer_v<T>)
inline void EventDispatcher(T value) {
    EventDispatcher<E>(static_cast<void *>(&value));
}
template<auto E>
inline void EventClear() {
    std::lock_guard lock(Event<E>::Mutex());
    Event<E>::Handlers().clear();
}
}// namespace Pupil
