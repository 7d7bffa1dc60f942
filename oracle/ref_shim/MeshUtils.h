// Shadows HoughForest/include/MeshUtils.h (which needs PCL, VTK and OpenCV -- none installed).  TEST INFRASTRUCTURE, see
// Eigen/Dense.  Same public types and the members HFTest.cpp calls.  icp() is NOT restated here: oracle/build_ref.py
// compiles the head of the reference's own MeshUtils::icp (the pre-ICP pose, MeshUtils.cpp:423-440) and its
// get_rotmat_from_yaw_pitch_roll (:29-59) from /root/reference and appends a call to record_icp(); the ICP refinement,
// hypothesis scoring and joint optimisation that follow are outside the hot path and are stubs.
#ifndef MESH_UTILS_H
#define MESH_UTILS_H
#include <cmath>
#include <omp.h>
#include <fstream>
#include <string>
#include <vector>

#include <Eigen/Dense>
#include <boost/unordered_map.hpp>
#include <cv.h>
#include <glog/logging.h>

struct Hf6dShimHypothesis {  // what the hot path hands to MeshUtils (HFTest.cpp:927-934)
    int obj_id, row, col;
    float z, yaw, pitch, roll, location_score, pose_score;
    float rotmat[16];
};
std::vector<Hf6dShimHypothesis>& hf6d_shim_hypotheses();

class MeshUtils {
    float fx_, fy_, cx_, cy_;
    int num_threads_;
    Eigen::Matrix4f get_rotmat_from_yaw_pitch_roll(float yaw, float pitch, float roll);
    void record_icp(int obj_id, int row, int col, float z, float yaw, float pitch, float roll, const Eigen::Matrix4f& rotmat);

  public:
    MeshUtils() : fx_(575), fy_(575), cx_(319.5f), cy_(239.5f), num_threads_(1) {}

    struct HypothesisEvaluation {
        float clutter_score, similarity_score, inliers_ratio, visibility_ratio, location_score, pose_score,
            ground_truth_error, final_score;
        HypothesisEvaluation()
            : clutter_score(0), similarity_score(0), inliers_ratio(0), visibility_ratio(0), location_score(0),
              pose_score(0), ground_truth_error(0), final_score(0) {}
    };
    struct ObjectHypothesis {
        int obj_id;
        Eigen::Matrix4f rotmat;
        HypothesisEvaluation eval;
        ObjectHypothesis() : obj_id(-1) {}
        ObjectHypothesis(int obj_id_, const Eigen::Matrix4f& rotmat_) : obj_id(obj_id_), rotmat(rotmat_) {}
    };

    void setIntrinsics(float fx, float fy, float cx, float cy) { fx_ = fx; fy_ = fy; cx_ = cx; cy_ = cy; }
    void setScene(const cv::Mat&, const cv::Mat&, float = 2.0f) {}
    void insertObjectFromPLY(const std::string&, int, std::string = "", bool = false, float = -1.0, int = -1) {}
    Eigen::Matrix4f getGroundTruthPose(int, int) { return Eigen::Matrix4f(); }
    cv::Mat getObjMask(int, Eigen::Matrix4f) { return cv::Mat(); }
    void setFinalScoreThreshold(float) {}
    void setClutterThreshold(float) {}
    void setInliersThreshold(float) {}
    void setReg(float, float, float, float, float) {}
    void setGroupReg(float, float) {}
    void searchSingleObjectInstance(bool) {}
    void searchSingleObjectInGroup(bool) {}
    void useColorSimilarity(bool) {}
    void setNumThreads(int n) { num_threads_ = n; }
    void setClusteringOptions(double, int, float, float, float) {}
    void renderObject(cv::Mat&, int, const Eigen::Matrix4f&, float) {}

    // defined by the TU oracle/build_ref.py generates from MeshUtils.cpp
    void icp(int obj_id, int row, int col, float z, float yaw, float pitch, float roll, Eigen::Matrix4f& rotmat, int iter);
    // stubs (ref_driver.cpp): every hypothesis is accepted, scores are recorded next to the icp() arguments
    bool evaluate_hypothesis(ObjectHypothesis& h, float location_score, float pose_score);
    std::vector<int> optimize_hypotheses(std::vector<ObjectHypothesis>& h);
};
#endif
