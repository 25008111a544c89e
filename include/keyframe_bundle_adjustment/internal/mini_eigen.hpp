// mini_eigen.hpp -- the small subset of Eigen that limo's BundleAdjusterKeyframes / Keyframe / LandmarkSelector API
// exposes to its callers (Isometry3d, Vector2d/3d, Matrix3d, Quaterniond, AngleAxisd), for builds where Eigen is not
// installed.  Semantics follow Eigen 3.3: Transform::translate/rotate post-multiply, Quaternion(Matrix3) and
// toRotationMatrix use Eigen's formulas (the pose convention of the reference depends on them, definitions.hpp:75-83,
// definitions.cpp:14-28), DenseBase::isApprox is the relative Frobenius test.
#pragma once
#if __has_include(<Eigen/Eigen>) && !defined(KBA_FORCE_MINI_EIGEN)
#include <Eigen/Eigen>
#else
#include <array>
#include <cmath>
#include <cstddef>

#include <limits>

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW

// Real Eigen leaves default-constructed fixed-size objects UNINITIALISED.  This subset only exposes spellings that exist
// in Eigen 3.3; to make sure the facade does not lean on anything else either, default construction fills with NaN when
// KBA_EIGEN_POISON_UNINIT is defined (the strict build leg of tests/test_cpp_facade.py): code that reads a
// default-constructed vector or matrix before writing it then fails its tests instead of silently seeing zeros.
#ifdef KBA_EIGEN_POISON_UNINIT
#define KBA_EIGEN_FILL std::numeric_limits<double>::quiet_NaN()
#else
#define KBA_EIGEN_FILL 0.0
#endif

namespace Eigen {

struct Vector2d {
    double v[2]{KBA_EIGEN_FILL, KBA_EIGEN_FILL};
    Vector2d() = default;
    Vector2d(double a, double b) : v{a, b} {}
    double& operator[](int i) { return v[i]; }
    double operator[](int i) const { return v[i]; }
    double x() const { return v[0]; }
    double y() const { return v[1]; }
    Vector2d operator-(const Vector2d& o) const { return {v[0] - o.v[0], v[1] - o.v[1]}; }
    Vector2d operator+(const Vector2d& o) const { return {v[0] + o.v[0], v[1] + o.v[1]}; }
    double norm() const { return std::sqrt(v[0] * v[0] + v[1] * v[1]); }
    static Vector2d Zero() { return {0., 0.}; }
};

struct RowVector3d {  // what Vector3d::transpose() yields: only usable as the right factor of an outer product
    double v[3];
};

struct Vector3d {
    double v[3]{KBA_EIGEN_FILL, KBA_EIGEN_FILL, KBA_EIGEN_FILL};
    Vector3d() = default;
    Vector3d(double a, double b, double c) : v{a, b, c} {}
    explicit Vector3d(const double* p) : v{p[0], p[1], p[2]} {}
    double& operator[](int i) { return v[i]; }
    double operator[](int i) const { return v[i]; }
    double& operator()(int i) { return v[i]; }
    double operator()(int i) const { return v[i]; }
    double x() const { return v[0]; }
    double y() const { return v[1]; }
    double z() const { return v[2]; }
    double* data() { return v; }
    const double* data() const { return v; }
    double dot(const Vector3d& o) const { return v[0] * o.v[0] + v[1] * o.v[1] + v[2] * o.v[2]; }
    Vector3d cross(const Vector3d& o) const {
        return {v[1] * o.v[2] - v[2] * o.v[1], v[2] * o.v[0] - v[0] * o.v[2], v[0] * o.v[1] - v[1] * o.v[0]};
    }
    double squaredNorm() const { return dot(*this); }
    double norm() const { return std::sqrt(squaredNorm()); }
    void normalize() { const double n = norm(); v[0] /= n; v[1] /= n; v[2] /= n; }
    Vector3d normalized() const { Vector3d r = *this; r.normalize(); return r; }
    Vector3d operator+(const Vector3d& o) const { return {v[0] + o.v[0], v[1] + o.v[1], v[2] + o.v[2]}; }
    Vector3d operator-(const Vector3d& o) const { return {v[0] - o.v[0], v[1] - o.v[1], v[2] - o.v[2]}; }
    Vector3d operator-() const { return {-v[0], -v[1], -v[2]}; }
    Vector3d operator*(double s) const { return {v[0] * s, v[1] * s, v[2] * s}; }
    Vector3d operator/(double s) const { return {v[0] / s, v[1] / s, v[2] / s}; }
    Vector3d& operator+=(const Vector3d& o) { v[0] += o.v[0]; v[1] += o.v[1]; v[2] += o.v[2]; return *this; }
    Vector3d& operator/=(double s) { v[0] /= s; v[1] /= s; v[2] /= s; return *this; }
    RowVector3d transpose() const { return {{v[0], v[1], v[2]}}; }
    static Vector3d Zero() { return {0., 0., 0.}; }
};
inline Vector3d operator*(double s, const Vector3d& a) { return a * s; }

struct Matrix3d {
    double m[9]{KBA_EIGEN_FILL, KBA_EIGEN_FILL, KBA_EIGEN_FILL, KBA_EIGEN_FILL, KBA_EIGEN_FILL,
                KBA_EIGEN_FILL, KBA_EIGEN_FILL, KBA_EIGEN_FILL, KBA_EIGEN_FILL};  // row-major
    double& operator()(int i, int j) { return m[3 * i + j]; }
    double operator()(int i, int j) const { return m[3 * i + j]; }
    static Matrix3d Zero() { Matrix3d r; for (int i = 0; i < 9; ++i) r.m[i] = 0.; return r; }
    static Matrix3d Identity() { Matrix3d r = Zero(); r.m[0] = r.m[4] = r.m[8] = 1; return r; }
    Matrix3d transpose() const { Matrix3d r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r(i, j) = (*this)(j, i); return r; }
    Matrix3d operator*(const Matrix3d& o) const {
        Matrix3d r;
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double s = 0; for (int k = 0; k < 3; ++k) s += (*this)(i, k) * o(k, j); r(i, j) = s; }
        return r;
    }
    Vector3d operator*(const Vector3d& a) const {
        return {m[0] * a[0] + m[1] * a[1] + m[2] * a[2], m[3] * a[0] + m[4] * a[1] + m[5] * a[2], m[6] * a[0] + m[7] * a[1] + m[8] * a[2]};
    }
    Matrix3d operator+(const Matrix3d& o) const { Matrix3d r; for (int i = 0; i < 9; ++i) r.m[i] = m[i] + o.m[i]; return r; }
    Matrix3d operator-(const Matrix3d& o) const { Matrix3d r; for (int i = 0; i < 9; ++i) r.m[i] = m[i] - o.m[i]; return r; }
    Matrix3d& operator+=(const Matrix3d& o) { for (int i = 0; i < 9; ++i) m[i] += o.m[i]; return *this; }
    double determinant() const {
        return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
    }
    Matrix3d inverse() const {
        const double d = determinant();
        Matrix3d r;
        r.m[0] = (m[4] * m[8] - m[5] * m[7]) / d; r.m[1] = (m[2] * m[7] - m[1] * m[8]) / d; r.m[2] = (m[1] * m[5] - m[2] * m[4]) / d;
        r.m[3] = (m[5] * m[6] - m[3] * m[8]) / d; r.m[4] = (m[0] * m[8] - m[2] * m[6]) / d; r.m[5] = (m[2] * m[3] - m[0] * m[5]) / d;
        r.m[6] = (m[3] * m[7] - m[4] * m[6]) / d; r.m[7] = (m[1] * m[6] - m[0] * m[7]) / d; r.m[8] = (m[0] * m[4] - m[1] * m[3]) / d;
        return r;
    }
};
inline Matrix3d operator*(const Vector3d& a, const RowVector3d& b) {  // a * b.transpose()
    Matrix3d r;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r(i, j) = a[i] * b.v[j];
    return r;
}

struct Quaterniond;
struct AngleAxisd {
    double angle_ = 0;
    Vector3d axis_{1, 0, 0};
    AngleAxisd() = default;
    AngleAxisd(double a, const Vector3d& ax) : angle_(a), axis_(ax) {}
    explicit AngleAxisd(const Quaterniond& q);
    double angle() const { return angle_; }
    const Vector3d& axis() const { return axis_; }
    Matrix3d toRotationMatrix() const {  // Eigen/src/Geometry/AngleAxis.h
        Matrix3d R;
        const double s = std::sin(angle_), c = std::cos(angle_);
        const Vector3d cc = axis_ * (1 - c);
        double t;
        t = cc.x() * axis_.y(); R(0, 1) = t - s * axis_.z(); R(1, 0) = t + s * axis_.z();
        t = cc.x() * axis_.z(); R(0, 2) = t + s * axis_.y(); R(2, 0) = t - s * axis_.y();
        t = cc.y() * axis_.z(); R(1, 2) = t - s * axis_.x(); R(2, 1) = t + s * axis_.x();
        R(0, 0) = cc.x() * axis_.x() + c; R(1, 1) = cc.y() * axis_.y() + c; R(2, 2) = cc.z() * axis_.z() + c;
        return R;
    }
};

struct Quaterniond {
    double w_ = 1, x_ = 0, y_ = 0, z_ = 0;
    Quaterniond() = default;
    Quaterniond(double w, double x, double y, double z) : w_(w), x_(x), y_(y), z_(z) {}
    explicit Quaterniond(const Matrix3d& mat) {  // Eigen/src/Geometry/Quaternion.h quaternionbase_assign_impl<3,3>
        double t = mat(0, 0) + mat(1, 1) + mat(2, 2);
        double c[3];
        if (t > 0) {
            t = std::sqrt(t + 1.0); w_ = 0.5 * t; t = 0.5 / t;
            x_ = (mat(2, 1) - mat(1, 2)) * t; y_ = (mat(0, 2) - mat(2, 0)) * t; z_ = (mat(1, 0) - mat(0, 1)) * t;
        } else {
            int i = 0;
            if (mat(1, 1) > mat(0, 0)) i = 1;
            if (mat(2, 2) > mat(i, i)) i = 2;
            const int j = (i + 1) % 3, k = (j + 1) % 3;
            t = std::sqrt(mat(i, i) - mat(j, j) - mat(k, k) + 1.0);
            c[i] = 0.5 * t; t = 0.5 / t;
            w_ = (mat(k, j) - mat(j, k)) * t; c[j] = (mat(j, i) + mat(i, j)) * t; c[k] = (mat(k, i) + mat(i, k)) * t;
            x_ = c[0]; y_ = c[1]; z_ = c[2];
        }
    }
    explicit Quaterniond(const AngleAxisd& aa) {
        const double h = 0.5 * aa.angle(), s = std::sin(h);
        w_ = std::cos(h); x_ = s * aa.axis().x(); y_ = s * aa.axis().y(); z_ = s * aa.axis().z();
    }
    double w() const { return w_; }
    double x() const { return x_; }
    double y() const { return y_; }
    double z() const { return z_; }
    Matrix3d toRotationMatrix() const {  // no normalisation, as in Eigen
        Matrix3d R;
        const double tx = 2 * x_, ty = 2 * y_, tz = 2 * z_, twx = tx * w_, twy = ty * w_, twz = tz * w_;
        const double txx = tx * x_, txy = ty * x_, txz = tz * x_, tyy = ty * y_, tyz = tz * y_, tzz = tz * z_;
        R(0, 0) = 1 - (tyy + tzz); R(0, 1) = txy - twz; R(0, 2) = txz + twy;
        R(1, 0) = txy + twz; R(1, 1) = 1 - (txx + tzz); R(1, 2) = tyz - twx;
        R(2, 0) = txz - twy; R(2, 1) = tyz + twx; R(2, 2) = 1 - (txx + tyy);
        return R;
    }
    Quaterniond inverse() const { const double n = w_ * w_ + x_ * x_ + y_ * y_ + z_ * z_; return {w_ / n, -x_ / n, -y_ / n, -z_ / n}; }
    Quaterniond operator*(const Quaterniond& b) const {
        return {w_ * b.w_ - x_ * b.x_ - y_ * b.y_ - z_ * b.z_, w_ * b.x_ + x_ * b.w_ + y_ * b.z_ - z_ * b.y_,
                w_ * b.y_ - x_ * b.z_ + y_ * b.w_ + z_ * b.x_, w_ * b.z_ + x_ * b.y_ - y_ * b.x_ + z_ * b.w_};
    }
};
inline AngleAxisd::AngleAxisd(const Quaterniond& q) {  // Eigen/src/Geometry/AngleAxis.h operator=(QuaternionBase)
    double n = std::sqrt(q.x() * q.x() + q.y() * q.y() + q.z() * q.z());
    if (n != 0.0) {
        angle_ = 2.0 * std::atan2(n, std::fabs(q.w()));
        if (q.w() < 0) n = -n;
        axis_ = Vector3d(q.x() / n, q.y() / n, q.z() / n);
    } else { angle_ = 0; axis_ = Vector3d(1, 0, 0); }
}
inline Matrix3d operator*(const AngleAxisd& a, const AngleAxisd& b) { return a.toRotationMatrix() * b.toRotationMatrix(); }

struct Matrix4d {
    double m[16]{0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    double operator()(int i, int j) const { return m[4 * i + j]; }
    double& operator()(int i, int j) { return m[4 * i + j]; }
    bool isApprox(const Matrix4d& o, double prec = 1e-12) const {
        double d = 0, a = 0, b = 0;
        for (int i = 0; i < 16; ++i) { d += (m[i] - o.m[i]) * (m[i] - o.m[i]); a += m[i] * m[i]; b += o.m[i] * o.m[i]; }
        return d <= prec * prec * (a < b ? a : b);
    }
};

// Transform<double, 3, Isometry>
struct Isometry3d {
    Matrix3d R = Matrix3d::Identity();   // (a default-constructed Eigen::Transform is uninitialised too; every use in the
    Vector3d t = Vector3d::Zero();       //  facade starts from Identity())
    static Isometry3d Identity() { return {}; }
    void setIdentity() { *this = Isometry3d(); }
    Isometry3d& translate(const Vector3d& v) { t += R * v; return *this; }
    Isometry3d& rotate(const Matrix3d& r) { R = R * r; return *this; }
    Isometry3d& rotate(const AngleAxisd& a) { return rotate(a.toRotationMatrix()); }
    Isometry3d& rotate(const Quaterniond& q) { return rotate(q.toRotationMatrix()); }
    Isometry3d inverse() const { Isometry3d r; r.R = R.transpose(); r.t = -(r.R * t); return r; }
    Isometry3d operator*(const Isometry3d& o) const { Isometry3d r; r.R = R * o.R; r.t = R * o.t + t; return r; }
    Vector3d operator*(const Vector3d& p) const { return R * p + t; }
    Vector3d& translation() { return t; }
    const Vector3d& translation() const { return t; }
    const Matrix3d& rotation() const { return R; }
    const Matrix3d& linear() const { return R; }
    Matrix3d& linear() { return R; }
    Matrix4d matrix() const {
        Matrix4d M;
        for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) M(i, j) = R(i, j); M(i, 3) = t[i]; }
        M(3, 3) = 1;
        return M;
    }
    bool isApprox(const Isometry3d& o, double prec = 1e-12) const { return matrix().isApprox(o.matrix(), prec); }
};

}  // namespace Eigen
#endif
