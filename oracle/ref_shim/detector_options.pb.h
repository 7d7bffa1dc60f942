// Stand-in for the protoc output of HoughForest/include/proto/detector_options.proto (needs libprotobuf headers, absent).
// TEST INFRASTRUCTURE, see Eigen/Dense.  Same accessors and defaults as detector_options.proto:3-70.
#ifndef HF6D_SHIM_DETECTOR_OPTIONS_PB_H
#define HF6D_SHIM_DETECTOR_OPTIONS_PB_H
#include <string>
#include <vector>

namespace DetectorOptions {

class ObjectOptions {
  public:
    std::string name_, mesh_file_;
    int instances_, icp_iterations_, max_location_hypotheses_;
    float nn_search_radius_;
    bool align_z_axis_, should_detect_;
    ObjectOptions() : instances_(1), icp_iterations_(60), max_location_hypotheses_(12), nn_search_radius_(0.01f),
                      align_z_axis_(false), should_detect_(true) {}
    const std::string& name() const { return name_; }
    const std::string& mesh_file() const { return mesh_file_; }
    int instances() const { return instances_; }
    float nn_search_radius() const { return nn_search_radius_; }
    int icp_iterations() const { return icp_iterations_; }
    bool align_z_axis() const { return align_z_axis_; }
    int max_location_hypotheses() const { return max_location_hypotheses_; }
    bool should_detect() const { return should_detect_; }
};

class Options {
  public:
    std::vector<ObjectOptions> objects_;
    std::string forest_folder_, caffe_definition_, caffe_weights_;
    int stride_, gpu_, num_threads_, batch_size_, cluster_min_points_;
    float max_depth_range_, fx_, fy_, cx_, cy_, distance_threshold_;
    bool are_objects_segmented_;
    Options() : stride_(4), gpu_(-1), num_threads_(4), batch_size_(100), cluster_min_points_(5), max_depth_range_(0.25f),
                fx_(575), fy_(575), cx_(319.5f), cy_(239.5f), distance_threshold_(1.5f), are_objects_segmented_(false) {}
    int object_options_size() const { return (int)objects_.size(); }
    const ObjectOptions& object_options(int i) const { return objects_[i]; }
    const std::string& forest_folder() const { return forest_folder_; }
    const std::string& caffe_definition() const { return caffe_definition_; }
    const std::string& caffe_weights() const { return caffe_weights_; }
    int stride() const { return stride_; }
    int gpu() const { return gpu_; }
    int num_threads() const { return num_threads_; }
    float max_depth_range_in_patch_in_m() const { return max_depth_range_; }
    int batch_size() const { return batch_size_; }
    float fx() const { return fx_; }
    float fy() const { return fy_; }
    float cx() const { return cx_; }
    float cy() const { return cy_; }
    bool search_single_object_instance() const { return false; }
    bool search_single_object_in_group() const { return false; }
    bool use_color_similarity() const { return true; }
    float similarity_coeff() const { return 10; }
    float inliers_coeff() const { return 2.5f; }
    float clutter_coeff() const { return 1.4f; }
    float location_score_coeff() const { return 1; }
    float pose_score_coeff() const { return 0.7f; }
    float group_total_explain_coeff() const { return 0.5f; }
    float group_common_explain_coeff() const { return 0.3f; }
    float inliers_threshold() const { return 0.6f; }
    float clutter_threshold() const { return 0.6f; }
    float final_score_threshold() const { return 10; }
    float cluster_eps_angle_threshold() const { return 0.05f; }
    int cluster_min_points() const { return cluster_min_points_; }
    float cluster_curvature_threshold() const { return 0.1f; }
    float cluster_tolerance_near() const { return 0.03f; }
    float cluster_tolerance_far() const { return 0.05f; }
    float distance_threshold() const { return distance_threshold_; }
    bool are_objects_segmented() const { return are_objects_segmented_; }
};

}  // namespace DetectorOptions
#endif
