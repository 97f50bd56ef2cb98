// ref_world.cc - the reference's own bodies of the Frame / KeyFrame / MapPoint / Converter methods that surround the ORB
// front-end, compiled over the stand-in declarations of dropin/shim/orbslam_world.h (TEST INFRASTRUCTURE).
// The *.inc files are produced at build time by extract_functions.py from /root/reference (see Makefile); nothing in them is
// written or edited here.  Static data members are defined as R/src/Frame.cc:19-26, KeyFrame.cc:45, MapPoint.cc:12-13 do.
#include <thread>
#include "orbslam_world.h"

namespace ORB_SLAM3
{
long unsigned int Frame::nNextId = 0;
bool Frame::mbInitialComputations = true;
float Frame::cx, Frame::cy, Frame::fx, Frame::fy, Frame::invfx, Frame::invfy;
float Frame::mnMinX, Frame::mnMinY, Frame::mnMaxX, Frame::mnMaxY;
float Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv;
long unsigned int KeyFrame::nNextId = 0;
long unsigned int MapPoint::nNextId = 0;
mutex MapPoint::mGlobalMutex;

#include "gen/Converter.inc"
#include "gen/Frame.inc"
#include "gen/KeyFrame.inc"
#include "gen/MapPoint.inc"
}  // namespace ORB_SLAM3
