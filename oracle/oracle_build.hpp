// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_core.hpp header).
//
// build_octree restated in memory: src/octree/generation.rs (all of it).  "Files" are NodeFile
// byte vectors; every write goes through the same encode and every read through the same decode
// as the reference's RawNodeWriter / RawNodeReader, in the same stream order, so node contents are
// byte-identical to what the reference would leave on disk (given identical float semantics).
#pragma once
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <set>
#include <stdexcept>
#include <thread>

#include <cstdio>
#include <unistd.h>

#include "oracle_core.hpp"

namespace orc {

struct BuildParams {
    double resolution = 0.001;
    Aabb bbox{};
    bool with_intensity = false;
    int64_t max_points_per_node = 100000;  // generation.rs:37 (exposed for small deep test trees)
};

// Minimal task pool standing in for the reference's rayon scope (generation.rs:312-323,164-188,368-378).
struct TaskPool {
    std::vector<std::thread> threads;
    std::deque<std::function<void()>> q;
    std::mutex m;
    std::condition_variable cv, done_cv;
    size_t pending = 0;
    bool stop = false;
    explicit TaskPool(int n) {
        for (int i = 0; i < n; ++i)
            threads.emplace_back([this] {
                for (;;) {
                    std::function<void()> f;
                    {
                        std::unique_lock<std::mutex> l(m);
                        cv.wait(l, [this] { return stop || !q.empty(); });
                        if (q.empty()) return;
                        f = std::move(q.front());
                        q.pop_front();
                    }
                    f();
                    std::lock_guard<std::mutex> l(m);
                    if (--pending == 0) done_cv.notify_all();
                }
            });
    }
    void spawn(std::function<void()> f) {
        {
            std::lock_guard<std::mutex> l(m);
            ++pending;
            q.push_back(std::move(f));
        }
        cv.notify_one();
    }
    void wait_all() {
        std::unique_lock<std::mutex> l(m);
        done_cv.wait(l, [this] { return pending == 0; });
    }
    ~TaskPool() {
        {
            std::lock_guard<std::mutex> l(m);
            stop = true;
        }
        cv.notify_all();
        for (auto& t : threads) t.join();
    }
};

struct Builder {
    BuildParams P;
    int num_threads = 1;
    TaskPool* pool = nullptr;
    Cube root;
    std::map<NodeId, NodeFile> files;
    std::mutex mu;
    std::vector<NodeId> leaves;
    // "Faithful" variant: like the reference, every node's .xyz / .rgb (/ .intensity) content lives in files of
    // `disk_dir` between the steps - written when a split / subsample step finishes a node (DataWriter, node_writer.rs:36-89),
    // read back when the next step consumes it (RawNodeReader, raw.rs:127-216), removed when the node is split
    // (generation.rs:104-108).  Empty: everything stays in memory (same arithmetic and order, no file round trips).  The
    // provenance column `src` is this oracle's own extension and never goes to disk.
    std::string disk_dir;

    static bool put_file(const std::string& path, const void* data, size_t n) {
        FILE* f = std::fopen(path.c_str(), "wb");
        if (!f) return false;
        const size_t w = n ? std::fwrite(data, 1, n, f) : 0;
        std::fclose(f);
        return w == n;
    }
    template <class V>
    static void get_file(const std::string& path, V& out, size_t elems) {
        out.resize(elems);
        FILE* f = std::fopen(path.c_str(), "rb");
        if (!f) throw std::runtime_error("oracle (faithful): cannot read " + path);
        const size_t r = elems ? std::fread(out.data(), sizeof(out[0]), elems, f) : 0;
        std::fclose(f);
        if (r != elems) throw std::runtime_error("oracle (faithful): short read " + path);
    }
    // A finished node leaves the writer: to disk in the faithful variant (zero-point nodes have no files, node_writer.rs:78-89).
    void spill(NodeId id, NodeFile& f) const {
        if (disk_dir.empty() || f.rgb.empty()) return;
        const std::string stem = disk_dir + "/" + id.to_string();
        if (!put_file(stem + ".xyz", f.xyz.data(), f.xyz.size()) || !put_file(stem + ".rgb", f.rgb.data(), f.rgb.size()) ||
            (P.with_intensity && !put_file(stem + ".intensity", f.intensity.data(), f.intensity.size() * 4)))
            throw std::runtime_error("oracle (faithful): cannot write " + stem);
        f.disk_points = f.rgb.size() / 3;
        std::vector<uint8_t>().swap(f.xyz);
        std::vector<uint8_t>().swap(f.rgb);
        std::vector<float>().swap(f.intensity);
    }
    // A step opens a node for reading: from disk in the faithful variant.
    void unspill(NodeId id, NodeFile& f, bool remove_files) const {
        if (disk_dir.empty() || f.disk_points == 0) return;
        const std::string stem = disk_dir + "/" + id.to_string();
        const size_t n = f.disk_points;
        get_file(stem + ".xyz", f.xyz, n * 3 * (size_t)bytes_per_coordinate(f.enc));
        get_file(stem + ".rgb", f.rgb, n * 3);
        if (P.with_intensity) get_file(stem + ".intensity", f.intensity, n);
        f.disk_points = 0;
        if (remove_files) {
            ::unlink((stem + ".xyz").c_str());
            ::unlink((stem + ".rgb").c_str());
            if (P.with_intensity) ::unlink((stem + ".intensity").c_str());
        }
    }

    NodeFile new_writer(NodeId id) const {  // generation.rs:39-56 (RawNodeWriter::from_data_provider)
        NodeFile f;
        f.cube = find_bounding_cube(id, root);
        f.enc = position_encoding(f.cube, P.resolution);
        return f;
    }

    bool should_split_node(NodeId id, int64_t num_points) const {  // generation.rs:128-150
        if (num_points <= P.max_points_per_node) return false;
        Cube c = find_bounding_cube(id, root);
        if (c.edge <= P.resolution) return false;
        return true;
    }

    // generation.rs:58-126.  `next(i)` yields the i-th point of the node's stream.
    template <class Stream>
    void split(NodeId node_id, size_t n, Stream&& next, std::vector<NodeId>& leaf_nodes,
               std::vector<NodeId>& split_nodes) {
        NodeFile children[8];
        bool present[8] = {false, false, false, false, false, false, false, false};
        Cube cube = find_bounding_cube(node_id, root);
        for (size_t i = 0; i < n; ++i) {
            Point pt = next(i);
            unsigned k = child_index_of(cube, pt.p);  // generation.rs:78-83
            if (!present[k]) {
                children[k] = new_writer(node_id.child(k));
                present[k] = true;
            }
            node_append(children[k], pt, P.with_intensity);  // generation.rs:84-101
        }
        for (unsigned k = 0; k < 8; ++k) {  // generation.rs:110-124
            if (!present[k]) continue;
            NodeId cid = node_id.child(k);
            int64_t num_written = (int64_t)(children[k].xyz.size() / bytes_per_coordinate(children[k].enc) / 3);
            bool sp = should_split_node(cid, num_written);
            spill(cid, children[k]);
            {
                std::lock_guard<std::mutex> g(mu);
                files[cid] = std::move(children[k]);
            }
            (sp ? split_nodes : leaf_nodes).push_back(cid);
        }
    }

    void split_node_recursive(NodeId id) {  // generation.rs:152-193 (child tasks)
        NodeFile f;
        {
            std::lock_guard<std::mutex> g(mu);
            f = std::move(files[id]);
            files.erase(id);  // generation.rs:104-108: the split node's own .xyz is removed
        }
        unspill(id, f, true);
        std::vector<NodeId> leaf_nodes, split_nodes;
        size_t n = (size_t)f.num_points();
        bool wi = P.with_intensity;
        split(id, n, [&](size_t i) { return node_read(f, i, wi); }, leaf_nodes, split_nodes);
        f = NodeFile();
        {
            std::lock_guard<std::mutex> g(mu);
            for (auto l : leaf_nodes) leaves.push_back(l);
        }
        for (auto c : split_nodes) pool->spawn([this, c] { split_node_recursive(c); });
    }

    // generation.rs:195-253
    void subsample_children_into(NodeId node_id, std::vector<std::pair<NodeId, int64_t>>& out) {
        NodeFile parent_writer = new_writer(node_id);
        for (unsigned i = 0; i < 8; ++i) {
            NodeId child_id = node_id.child(i);
            NodeFile child;
            {
                std::lock_guard<std::mutex> g(mu);
                auto it = files.find(child_id);
                if (it == files.end()) continue;  // NodeNotFound -> continue (generation.rs:207-211)
                child = std::move(it->second);
            }
            unspill(child_id, child, true);  // the child is rewritten below (generation.rs:216-238)
            size_t n = (size_t)child.num_points();
            NodeFile child_writer = new_writer(child_id);
            for (size_t j = 0; j < n; ++j) {  // generation.rs:220-238
                Point pt = node_read(child, j, P.with_intensity);
                if (j % 8 == 0)
                    node_append(parent_writer, pt, P.with_intensity);
                else
                    node_append(child_writer, pt, P.with_intensity);
            }
            int64_t nw = (int64_t)(child_writer.xyz.size() / bytes_per_coordinate(child_writer.enc) / 3);
            spill(child_id, child_writer);
            {
                std::lock_guard<std::mutex> g(mu);
                if (nw == 0)
                    files.erase(child_id);  // node_writer.rs:78-89: empty files are deleted on drop
                else
                    files[child_id] = std::move(child_writer);
            }
            out.push_back({child_id, nw});  // generation.rs:241-243
        }
        int64_t pn = (int64_t)(parent_writer.xyz.size() / bytes_per_coordinate(parent_writer.enc) / 3);
        if (node_id.level() == 0) out.push_back({node_id, pn});  // generation.rs:246-251
        spill(node_id, parent_writer);
        std::lock_guard<std::mutex> g(mu);
        if (pn == 0)
            files.erase(node_id);
        else
            files[node_id] = std::move(parent_writer);
    }

    // generation.rs:289-403.  x/y/z are read with element stride `stride` (1 = SoA, 3 = AoS xyz).
    Octree build(size_t n, const double* x, const double* y, const double* z, size_t stride,
                 const uint8_t* rgb, const float* intensity) {
        root = Cube::bounding(P.bbox);
        Octree oct;
        oct.resolution = P.resolution;
        oct.bbox = P.bbox;
        oct.with_intensity = P.with_intensity;

        // Root split: the caller's stream, consumed serially (generation.rs:312-323).
        std::vector<NodeId> leaf_nodes, split_nodes;
        bool wi = P.with_intensity;
        split(
            NodeId(), n,
            [&](size_t i) {
                Point pt;
                pt.p = {x[i * stride], y[i * stride], z[i * stride]};
                pt.rgb[0] = rgb[3 * i];
                pt.rgb[1] = rgb[3 * i + 1];
                pt.rgb[2] = rgb[3 * i + 2];
                pt.intensity = wi ? intensity[i] : 0.f;
                pt.src = i;
                return pt;
            },
            leaf_nodes, split_nodes);
        for (auto l : leaf_nodes) leaves.push_back(l);
        TaskPool tp(num_threads);
        pool = &tp;
        for (auto c : split_nodes) tp.spawn([this, c] { split_node_recursive(c); });
        tp.wait_all();

        std::vector<NodeId> nodes_to_subsample = leaves;  // generation.rs:325-330
        uint8_t deepest_level = 0;
        for (auto id : nodes_to_subsample) deepest_level = std::max(deepest_level, id.level());
        std::map<NodeId, int64_t> finished_nodes;

        for (int current_level = deepest_level; current_level >= 1; --current_level) {  // :335-387
            std::vector<NodeId> rest;
            std::set<NodeId> parent_ids;
            for (auto id : nodes_to_subsample) {
                if (id.level() == current_level)
                    parent_ids.insert(id.parent());
                else
                    rest.push_back(id);
            }
            nodes_to_subsample.swap(rest);
            std::vector<NodeId> parents(parent_ids.begin(), parent_ids.end());
            std::vector<std::vector<std::pair<NodeId, int64_t>>> outs(parents.size());
            for (size_t pi = 0; pi < parents.size(); ++pi)
                tp.spawn([this, pi, &parents, &outs] { subsample_children_into(parents[pi], outs[pi]); });
            tp.wait_all();
            for (auto& o : outs)
                for (auto& kv : o) finished_nodes[kv.first] = kv.second;
            for (auto p : parents) nodes_to_subsample.push_back(p);
        }

        for (auto& kv : finished_nodes) {  // generation.rs:389-397
            NodeMeta m;
            m.num_points = kv.second;
            m.cube = find_bounding_cube(kv.first, root);
            m.enc = position_encoding(m.cube, P.resolution);
            oct.nodes[kv.first] = m;
        }
        oct.files = std::move(files);
        return oct;
    }
};

// find_bounding_box, generation.rs:256-270 (Aabb::new(pos,pos) then grow; empty -> zero box)
inline Aabb find_bounding_box(size_t n, const double* x, const double* y, const double* z, size_t stride) {
    if (n == 0) return Aabb{{0, 0, 0}, {0, 0, 0}};
    Vec3 p0{x[0], y[0], z[0]};
    Aabb b = Aabb::make(p0, p0);
    for (size_t i = 0; i < n; ++i) b.grow({x[i * stride], y[i * stride], z[i * stride]});
    return b;
}

}  // namespace orc
