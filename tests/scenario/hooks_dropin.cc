// hooks_dropin.cc - TEST INFRASTRUCTURE: nothing to prepare around an extraction in the drop-in builds.
void scenario_before_extract() {}
void scenario_after_extract() {}
