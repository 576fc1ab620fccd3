extern "C" int hlala_ref_dummy() { return 0; }
