// bundle_adjuster_keyframes.hpp -- source-compatible facade of limo's BundleAdjusterKeyframes / Keyframe / LandmarkSelector
// API (reference: keyframe_bundle_adjustment/include/keyframe_bundle_adjustment/{bundle_adjuster_keyframes,keyframe,
// landmark_selector}.hpp, internal/definitions.hpp, matches_msg_types/*.hpp) whose solve() / adjustPoseOnly() run on
// the B200 through the C ABI of kba_b200.h instead of Ceres.  Same namespaces, class names, public members, method
// signatures and exceptions; the private ceres::Problem member is replaced by a kba_handle.
#pragma once
#include <array>
#include <cstdint>
#include <exception>
#include <limits>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <vector>

#include "internal/mini_eigen.hpp"

struct kba_handle;

namespace matches_msg_types {
using TimestampNSec = uint64_t;
struct FeaturePoint {  // feature_point.hpp:4-40
    FeaturePoint() {}
    FeaturePoint(float u, float v) : u(u), v(v), d(-1) {}
    FeaturePoint(float u, float v, float d) : u(u), v(v), d(d) {}
    float u, v, d;
};
struct Tracklet {  // tracklet.hpp
    std::vector<FeaturePoint> feature_points;
    unsigned long id;
    unsigned long age;
    bool is_outlier{false};
    int label{-2};
};
struct Tracklets {  // tracklets.hpp
    std::vector<TimestampNSec> stamps;
    std::vector<Tracklet> tracks;
};
}  // namespace matches_msg_types

namespace keyframe_bundle_adjustment {

// ---- internal/definitions.hpp:13-34 ----
using CameraId = unsigned long;
using TimestampNSec = matches_msg_types::TimestampNSec;
using TimestampSec = double;
using LandmarkId = unsigned long;
using KeyframeId = unsigned long;
using CameraIds = std::vector<CameraId>;
using PoseId = KeyframeId;
using EigenPose = Eigen::Isometry3d;
using Pose = std::array<double, 7>;
using ResidualId = long;  // opaque (ceres::ResidualBlockId in the reference)
using Direction = std::array<double, 3>;
using FeaturePoint = matches_msg_types::FeaturePoint;
using Tracklet = matches_msg_types::Tracklet;
using Tracklets = matches_msg_types::Tracklets;
using Measurement = FeaturePoint;

struct Plane {  // definitions.hpp:27-34
    Plane() : direction{{0., 0., 1.}}, distance(-std::numeric_limits<double>::max()) {}
    Direction direction;
    double distance;
};

struct Landmark {  // definitions.hpp:42-68
    using Ptr = std::shared_ptr<Landmark>;
    using ConstPtr = std::shared_ptr<const Landmark>;
    Landmark() {}
    Landmark(const Eigen::Vector3d& p, bool has_depth = false) : pos{{p[0], p[1], p[2]}}, has_measured_depth(has_depth) {}
    std::array<double, 3> pos;
    bool has_measured_depth{false};
    bool is_ground_plane{false};
    double weight{1.};
};

Pose convert(EigenPose p);                    // definitions.cpp:14-28
EigenPose convert(const Pose& pose);          // definitions.hpp:75-88
TimestampSec convert(const TimestampNSec& ts);
TimestampNSec convert(const TimestampSec& ts);
double calcQuaternionDiff(const Pose& p0, const Pose& p1);  // definitions.cpp:104-111

struct Camera {  // definitions.hpp:93-124
    using Ptr = std::shared_ptr<Camera>;
    Camera(double f, const Eigen::Vector2d& pp, const EigenPose& pose_cam_veh);
    Eigen::Matrix3d getIntrinsicMatrix() const;
    EigenPose getEigenPose() const;
    double focal_length;
    Eigen::Vector2d principal_point;
    Pose pose_camera_vehicle;
    Eigen::Matrix3d intrin_inv;
};

class Keyframe {  // keyframe.hpp:27-196
public:
    enum class FixationStatus { Pose, Scale, None };
    using Ptr = std::shared_ptr<Keyframe>;
    using ConstPtr = std::shared_ptr<const Keyframe>;
    Keyframe() {}
    Keyframe(TimestampNSec timestamp, const Tracklets& tracklets, std::map<CameraId, Camera::Ptr> cameras,
             std::map<LandmarkId, CameraIds> landmark_to_cameras, EigenPose p,
             FixationStatus fix_stat = FixationStatus::None, Plane ground_plane = Plane());
    Keyframe(TimestampNSec timestamp, const Tracklets& tracklets, Camera::Ptr camera, EigenPose p,
             FixationStatus fix_stat = FixationStatus::None, Plane ground_plane = Plane());
    bool operator<(const Keyframe& kf) const { return timestamp_ < kf.timestamp_; }
    void assignMeasurements(const Tracklets&, const CameraId&);
    void assignMeasurements(const Tracklets& tracklets, const std::map<LandmarkId, CameraIds>& landmark_lookup);
    void assignPose(const EigenPose& p) { pose_ = convert(p); }
    Measurement& getMeasurement(LandmarkId lm_id, CameraId cam_id) { return measurements_.at(lm_id).at(cam_id); }
    const Measurement& getMeasurement(LandmarkId lm_id, CameraId cam_id) const { return measurements_.at(lm_id).at(cam_id); }
    std::map<CameraId, Measurement> getMeasurements(LandmarkId lm_id) const;
    bool hasMeasurement(const LandmarkId& lm_id, const CameraId& cam_id) const;
    bool hasMeasurement(LandmarkId lm_id) const;
    std::map<CameraId, Eigen::Vector3d> getProjectedLandmarkPosition(const std::pair<LandmarkId, Landmark::ConstPtr>& landmark_origin) const;
    EigenPose getEigenPose() const { return convert(pose_); }
    std::shared_ptr<Pose> getPosePtr() const { return std::make_shared<Pose>(pose_); }

    TimestampNSec timestamp_;
    std::map<CameraId, Camera::Ptr> cameras_;
    FixationStatus fixation_status_;
    Pose pose_;
    Plane local_ground_plane_;
    std::map<LandmarkId, std::map<CameraId, Measurement>> measurements_;
    bool is_active_;
};

// ---- landmark selection (landmark_selection_scheme_base.hpp, landmark_selector.hpp) ----
class LandmarkSchemeBase {
public:
    using LandmarkMap = std::map<LandmarkId, Landmark::ConstPtr>;
    using KeyframeMap = std::map<KeyframeId, Keyframe::ConstPtr>;
    virtual ~LandmarkSchemeBase() = default;
    virtual std::set<LandmarkId> getSelection(const LandmarkMap& landmarks, const KeyframeMap& keyframes) const = 0;
    std::string identifier = "";
};
class LandmarkRejectionSchemeBase : public LandmarkSchemeBase {
public:
    using Ptr = std::shared_ptr<LandmarkRejectionSchemeBase>;
    using ConstPtr = std::shared_ptr<const LandmarkRejectionSchemeBase>;
};
class LandmarkSelectionSchemeBase : public LandmarkSchemeBase {
public:
    using Ptr = std::shared_ptr<LandmarkSelectionSchemeBase>;
    using ConstPtr = std::shared_ptr<const LandmarkSelectionSchemeBase>;
};
class LandmarkSparsificationSchemeBase : public LandmarkSchemeBase {
public:
    using Ptr = std::shared_ptr<LandmarkSparsificationSchemeBase>;
    using ConstPtr = std::shared_ptr<const LandmarkSparsificationSchemeBase>;
};
class LandmarkRejectionSchemeCheirality : public LandmarkRejectionSchemeBase {  // landmark_selection_scheme_cheirality.cpp:22-60
public:
    std::set<LandmarkId> getSelection(const LandmarkMap& landmarks, const KeyframeMap& keyframes) const override;
    static ConstPtr createConst() { return ConstPtr(new LandmarkRejectionSchemeCheirality()); }
    static Ptr create() { return Ptr(new LandmarkRejectionSchemeCheirality()); }
};

class LandmarkSelector {  // landmark_selector.hpp:40-345
public:
    virtual ~LandmarkSelector() = default;
    void addScheme(LandmarkSelectionSchemeBase::ConstPtr scheme) { selection_schemes_.push_back(scheme); }
    void addScheme(LandmarkSparsificationSchemeBase::ConstPtr scheme) { sparsification_schemes_.push_back(scheme); }
    void addScheme(LandmarkRejectionSchemeBase::ConstPtr scheme) { rejection_schemes_.push_back(scheme); }
    std::set<LandmarkId> select(const std::map<LandmarkId, Landmark::ConstPtr>& landmarks,
                                const std::map<KeyframeId, Keyframe::ConstPtr>& kfs);
    const std::map<LandmarkId, unsigned int>& getUnselectedLandmarks() const { return unselected_lms_; }
    std::set<LandmarkId> getLastSelection() const { return last_selected_lms_; }
    void clearOutliers() { outlier_ids_.clear(); }
    const std::set<LandmarkId>& getOutliers() const { return outlier_ids_; }
    void setOutlier(LandmarkId id) { outlier_ids_.insert(id); }
    void setOutlier(const std::set<LandmarkId>& ids) { for (const auto& el : ids) setOutlier(el); }

    std::vector<LandmarkSelectionSchemeBase::ConstPtr> selection_schemes_;
    std::vector<LandmarkSparsificationSchemeBase::ConstPtr> sparsification_schemes_;
    std::vector<LandmarkRejectionSchemeBase::ConstPtr> rejection_schemes_;
    std::set<LandmarkId> outlier_ids_;

private:
    std::map<LandmarkId, unsigned int> unselected_lms_;
    std::map<LandmarkId, TimestampNSec> last_time_seen_;
    std::set<LandmarkId> last_selected_lms_;
};

// ---- the adjuster (bundle_adjuster_keyframes.hpp:40-335) ----
class BundleAdjusterKeyframes {
public:
    using UPtr = std::unique_ptr<BundleAdjusterKeyframes>;
    using Ptr = std::shared_ptr<BundleAdjusterKeyframes>;
    using v3 = Eigen::Vector3d;

    struct NotEnoughKeyframesException : public std::exception {
        NotEnoughKeyframesException(size_t num_is, size_t num_should_be);
        const char* what() const noexcept override { return msg.c_str(); }
        size_t num_is, num_should_be;
        std::string msg;
    };
    struct KeyframeNotFoundException : public std::exception {
        explicit KeyframeNotFoundException(TimestampNSec timestamp);
        const char* what() const noexcept override { return msg.c_str(); }
        TimestampNSec ts_;
        std::string msg;
    };
    struct OutlierRejectionOptions {  // :79-89
        double depth_thres{0.16};
        double reprojection_thres{1.6};
        double depth_quantile{0.95};
        double reprojection_quantile{0.95};
        int num_iterations{1};
    };

    BundleAdjusterKeyframes();
    ~BundleAdjusterKeyframes();
    BundleAdjusterKeyframes(const BundleAdjusterKeyframes&) = delete;
    BundleAdjusterKeyframes& operator=(const BundleAdjusterKeyframes&) = delete;

    void push(const Keyframe& kf);
    void push(const std::vector<Keyframe>& kfs);
    std::string solve();
    void deactivateKeyframes(int min_num_connecting_landmarks = 3, int min_size_optimization_window = 4,
                             int max_size_optimization_window = 20);
    const Keyframe& getKeyframe(TimestampSec timestamp = -1.) const;
    std::map<KeyframeId, Keyframe::Ptr> getActiveKeyframePtrs() const;
    std::map<KeyframeId, Keyframe::ConstPtr> getActiveKeyframeConstPtrs() const;
    std::vector<Keyframe::Ptr> getSortedActiveKeyframePtrs() const;
    std::vector<std::pair<KeyframeId, Keyframe::Ptr>> getSortedIdsWithActiveKeyframePtrs() const;
    std::map<LandmarkId, Landmark::ConstPtr> getActiveLandmarkConstPtrs() const;
    std::map<LandmarkId, Landmark::ConstPtr> getSelectedLandmarkConstPtrs() const;
    std::string adjustPoseOnly(Keyframe&);
    bool calculateLandmark(const Keyframe& kf, const LandmarkId& lId, v3& posAbs);
    bool calculateLandmark(const LandmarkId& lId, v3& posAbs);
    void set_solver_time(double solver_time_sec) { this->solver_time_sec = solver_time_sec; }
    void updateLabels(const Tracklets& t, double shrubbery_weight = 1.);

    std::map<KeyframeId, Keyframe::Ptr> keyframes_;
    std::map<LandmarkId, Landmark::Ptr> landmarks_;
    std::set<KeyframeId> active_keyframe_ids_;
    std::set<LandmarkId> active_landmark_ids_;
    std::set<LandmarkId> selected_landmark_ids_;
    OutlierRejectionOptions outlier_rejection_options_;
    std::unique_ptr<LandmarkSelector> landmark_selector_;
    std::map<std::string, std::set<int>> labels_{{"outliers", {23, 24, 25, 26, 27, 28, 29, 30, 31, 32, 33}},
                                                 {"shrubbery", {21}},
                                                 {"ground", {6, 7, 8, 9, 10}}};

private:
    std::map<LandmarkId, Landmark::ConstPtr> filterLandmarksById(const std::set<LandmarkId>& ids) const;
    std::string runWindow(const std::vector<Keyframe*>& kfs, const std::vector<LandmarkId>& lm_ids, bool motion_only,
                          Keyframe* speed_prior_for);
    kba_handle* handle_{nullptr};  // replaces std::shared_ptr<ceres::Problem> problem_
    double solver_time_sec;
};

// Triangulator::triangulate_rays (internal/triangulator.hpp:51-75); poses are origin <- camera
Eigen::Vector3d triangulate_rays(const std::vector<std::pair<EigenPose, Eigen::Vector3d>>& poses_rays);

}  // namespace keyframe_bundle_adjustment
