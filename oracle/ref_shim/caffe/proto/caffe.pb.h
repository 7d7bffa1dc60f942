#include <caffe/caffe.hpp>
