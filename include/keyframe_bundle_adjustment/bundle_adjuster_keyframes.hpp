// bundle_adjuster_keyframes.hpp -- source-compatible facade of limo's BundleAdjusterKeyframes (reference:
// keyframe_bundle_adjustment/include/keyframe_bundle_adjustment/bundle_adjuster_keyframes.hpp:40-335) whose solve() /
// adjustPoseOnly() run on the B200 through the C ABI of kba_b200.h instead of Ceres.  Same namespace, class name, public
// members, method signatures and exceptions; the private ceres::Problem member is replaced by a kba_handle.  The types
// around it live where the reference keeps them: internal/definitions.hpp, keyframe.hpp, landmark_selector.hpp,
// landmark_selection_schemes.hpp, matches_msg_types/*.hpp.
#pragma once
#include <array>
#include <exception>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <vector>

#include "internal/triangulator.hpp"
#include "keyframe.hpp"
#include "landmark_selector.hpp"

struct kba_handle;
struct kba_track;

namespace keyframe_bundle_adjustment {

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
    // Not in the reference: keep the pushed keyframes on the device (kba_track_*, include/kba_b200.h) so that solve() sends only
    // the lists of active keyframes / selected landmarks instead of re-packing and re-uploading the window (the reference
    // rebuilds its ceres::Problem per call, cpp:635-637).  On by default; windows the device-resident store cannot take (more than
    // 30 keyframes, ground-plane residuals, a camera that was not there at the first push) fall back to the rebuild path.
    void set_persistent_window(bool on) { persistent_window_ = on; }
    // host -> device bytes of the last solve() (either path) and of all push() calls so far (persistent path)
    long long lastSolveUploadBytes() const { return last_solve_h2d_; }
    long long pushUploadBytes() const { return push_h2d_; }
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
    // persistent device-resident window
    bool ensureHandle();
    bool trackPush(const Keyframe& kf);
    bool solveTracked(const std::vector<Keyframe*>& kfs, const std::vector<LandmarkId>& lm_ids, std::string& report);
    kba_track* track_{nullptr};
    bool persistent_window_{true}, track_failed_{false};
    std::map<KeyframeId, int> kf_slot_;
    std::map<LandmarkId, int> lm_slot_;
    std::vector<int> free_kf_slots_;
    std::vector<std::array<double, 10>> track_cams_;  // camera values (f, pp, pose_camera_vehicle) the track was created with
    std::set<LandmarkId> new_landmarks_, dirty_weights_;
    long long last_solve_h2d_{0}, push_h2d_{0};
};

}  // namespace keyframe_bundle_adjustment
