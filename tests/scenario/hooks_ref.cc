// hooks_ref.cc - TEST INFRASTRUCTURE: scenario_ref runs the reference's ORBextractor under the monotonic node arena
// (oracle/ref/ref_alloc.cc: address order of std::list<ExtractorNode> nodes = creation order, the canonical tie rule).
#include <list>
#include <opencv2/core/core.hpp>
#include "ORBextractor.h"
#include "../../oracle/ref/ref_alloc.h"
void scenario_before_extract() { ref_arena_begin(sizeof(std::_List_node<ORB_SLAM3::ExtractorNode>)); }
void scenario_after_extract() { ref_arena_end(); }
