// query_api.inl — placeholder
extern "C" {
int pcv_nodes_in_location(const pcv_octree*, const pcv_location*, uint64_t*, uint64_t, uint64_t*) { return fail(PCV_ERR_UNSUPPORTED, "nyi"); }
int pcv_visible_nodes(const pcv_octree*, const double*, uint64_t*, uint64_t, uint64_t*) { return fail(PCV_ERR_UNSUPPORTED, "nyi"); }
int pcv_query_points(const pcv_octree*, const pcv_location*, const pcv_interval*, uint32_t, uint64_t, pcv_batch_cb, void*) { return fail(PCV_ERR_UNSUPPORTED, "nyi"); }
int pcv_query_batch_device(const pcv_octree*, const pcv_location*, uint32_t, const pcv_interval*, uint32_t, uint64_t*, uint64_t*) { return fail(PCV_ERR_UNSUPPORTED, "nyi"); }
int pcv_xray_tile(const pcv_octree*, const double*, const double*, uint32_t, uint32_t, const double*, uint8_t*, uint32_t*, int*) { return fail(PCV_ERR_UNSUPPORTED, "nyi"); }
int pcv_prefix_histogram_device(pcv_ctx*, const pcv_points*, double, const double*, const double*, uint32_t, uint64_t*) { return fail(PCV_ERR_UNSUPPORTED, "nyi"); }
int pcv_prefix_pack_device(pcv_ctx*, const pcv_points*, const uint64_t*, double, const double*, const double*, uint32_t, const int32_t*, uint32_t, double*, uint8_t*, float*, uint64_t*, uint64_t*) { return fail(PCV_ERR_UNSUPPORTED, "nyi"); }
}
