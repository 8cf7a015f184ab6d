// C++ restatement of the reference's own integration tests for this path (src/octree/tests.rs:18-136), written against
// include/pcv.hpp so that it reads like the original: build the 100 001-point octree, stream it through the
// ParallelIterator with an erroring consumer and with a large batch.
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <string>

#include "../../include/pcv.hpp"

static const size_t NUM_POINTS = 100001;

#define ASSERT_EQ(a, b)                                                                         \
    do {                                                                                        \
        if (!((a) == (b))) {                                                                    \
            fprintf(stderr, "%s:%d: %s != %s (%lld vs %lld)\n", __FILE__, __LINE__, #a, #b, (long long)(a), (long long)(b)); \
            exit(1);                                                                            \
        }                                                                                       \
    } while (0)

static pcv::Octree build_test_octree(pcv::Context& ctx, const std::string& dir) {  // tests.rs:18-46
    pcv::PointsBatch batch;
    batch.position.assign(NUM_POINTS, {0.0, 0.0, 0.0});
    batch.color.assign(NUM_POINTS, {255, 0, 0});
    batch.position[NUM_POINTS - 1] = {-200., -40., 30.};
    pcv::Aabb bounding_box(batch.position[0], batch.position[NUM_POINTS - 1]);
    std::vector<pcv::PointsBatch> input{batch};
    pcv::build_octree(ctx, dir, 1.0, bounding_box, input.begin(), input.end(), {"color"});
    return pcv::Octree::from_directory(ctx, dir);  // Octree::from_data_provider(OnDiskDataProvider{directory})
}

struct Consumer {  // tests.rs:48-81
    size_t max_num_points, num_received_points = 0, num_received_callbacks = 0;
    explicit Consumer(size_t m) : max_num_points(m) {}
    bool consume(pcv::PointsBatch&& b) {
        num_received_callbacks += 1;
        num_received_points += b.position.size();
        return num_received_points < max_num_points;  // false == Err("Maximum number of points reached")
    }
};

int main(int argc, char** argv) {
    const std::string dir = argc > 1 ? argv[1] : "/tmp/pcv_cpp_octree";
    pcv::Context ctx(0);

    {  // test_batch_iterator (tests.rs:83-112)
        const size_t batch_size = 5000, max_num_points = 13000;
        Consumer c(max_num_points);
        pcv::Octree octree = build_test_octree(ctx, dir);
        pcv::PointQuery location;  // attributes: ["color"], location: AllPoints
        pcv::ParallelIterator it({&octree}, location, batch_size, 3, 4);
        const bool ok = it.try_for_each_batch([&](pcv::PointsBatch&& b) { return c.consume(std::move(b)); });
        ASSERT_EQ(ok, false);  // expect_err: the iterator must error when the callback errors
        if (!(c.num_received_points >= c.max_num_points)) return 1;
        ASSERT_EQ(c.num_received_callbacks, 3u);          // "The number of points doesn't fit in two batches"
        ASSERT_EQ(c.num_received_points, 3 * batch_size);  // "The callback received full batches"
    }
    {  // test_batch_iterator_more_points (tests.rs:114-136)
        const size_t batch_size = NUM_POINTS / 2, max_num_points = NUM_POINTS + 30000;
        Consumer c(max_num_points);
        pcv::Octree octree = build_test_octree(ctx, dir);
        pcv::PointQuery location;
        pcv::ParallelIterator it({&octree}, location, batch_size, 2, 2);
        const bool ok = it.try_for_each_batch([&](pcv::PointsBatch&& b) { return c.consume(std::move(b)); });
        ASSERT_EQ(ok, true);
        ASSERT_EQ(c.num_received_points, NUM_POINTS);
    }
    {  // the hand-derivable golden tree of the same scenario + node ids (node.rs Display)
        pcv::Octree octree = build_test_octree(ctx, dir);
        ASSERT_EQ(octree.nodes().size(), 3u);
        ASSERT_EQ(octree.num_points(), (int64_t)NUM_POINTS);
        for (auto& m : octree.nodes()) {
            pcv::NodeId id{m.id_high, m.id_low};
            const std::string s = id.to_string();
            if (s == "r") ASSERT_EQ(m.num_points, 12501);
            else if (s == "r0") ASSERT_EQ(m.num_points, 0);
            else if (s == "r4") ASSERT_EQ(m.num_points, 87500);
            else return 2;
        }
        pcv::NodeData d = octree.get_node_data(pcv::NodeId{0, 0});
        ASSERT_EQ(d.position.size(), 12501u * 3u);  // Uint8 encoding
        ASSERT_EQ(d.color.size(), 12501u * 3u);
    }
    {  // PLY input + /nodes_data blob + colour tile through the C++ mirror
        const std::string ply = dir + "_in.ply";
        const uint32_t n = 20000;
        {
            FILE* f = fopen(ply.c_str(), "wb");
            if (!f) return 3;
            fprintf(f, "ply\nformat binary_little_endian 1.0\ncomment offset: 10 20 30\nelement vertex %u\nproperty float x\nproperty float y\n"
                       "property float z\nproperty uchar red\nproperty uchar green\nproperty uchar blue\nend_header\n", n);
            for (uint32_t i = 0; i < n; ++i) {
                const float xyz[3] = {(float)(i % 200) * 0.5f, (float)((i / 200) % 100) * 0.5f, (float)(i % 7)};
                const uint8_t rgb[3] = {(uint8_t)(i % 251), 17, (uint8_t)(i % 3)};
                fwrite(xyz, 4, 3, f);
                fwrite(rgb, 1, 3, f);
            }
            fclose(f);
        }
        pcv::Octree t = pcv::build_octree_from_file(ctx, "", 0.01, ply);
        ASSERT_EQ(t.num_points(), (int64_t)n);
        std::vector<pcv::NodeId> ids;
        size_t expect = 0;
        for (auto& m : t.nodes())
            if (m.num_points > 0) {
                ids.push_back(pcv::NodeId{m.id_high, m.id_low});
                const size_t bpc = m.position_encoding == 1 ? 1 : m.position_encoding == 2 ? 2 : m.position_encoding == 3 ? 4 : 8;
                expect += 40 + (((size_t)m.num_points * 3 * bpc + 7) / 8) * 8 + (((size_t)m.num_points * 3 + 7) / 8) * 8;
            }
        const std::vector<uint8_t> blob = t.nodes_data_blob(ids);
        ASSERT_EQ(blob.size(), expect);
        uint32_t n0 = 0;
        memcpy(&n0, blob.data() + 32, 4);
        for (auto& m : t.nodes())
            if (m.num_points > 0) {  // the blob starts with the header of the first requested node
                ASSERT_EQ((int64_t)n0, m.num_points);
                break;
            }
        std::vector<uint8_t> rgba;
        const bool any = t.xray_tile_attr(pcv::Aabb({10, 20, 30}, {110, 70, 37}), 64, 32, PCV_XRAY_COLORED, 0.f, 0.f, 0, rgba);
        ASSERT_EQ(any, true);
        size_t covered = 0;
        for (size_t px = 0; px < rgba.size() / 4; ++px) covered += rgba[px * 4 + 3] == 255;
        if (covered < 100) return 4;
        remove(ply.c_str());
    }
    printf("cpp octree tests OK\n");
    return 0;
}
