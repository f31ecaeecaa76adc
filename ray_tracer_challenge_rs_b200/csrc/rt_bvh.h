// rt_bvh.h — host-side BVH builder over the world-space boxes of the bounded shapes.
//
// Not part of the reference (it tests every shape for every ray, composites/world.rs:31-33).  The
// hierarchy only decides WHICH exact ray-shape tests run; every test that can produce an
// intersection the query cares about still runs, in the reference's arithmetic, so results are
// unchanged (boxes are inflated far beyond f64 rounding, and hit selection is by (distance, world
// order), independent of visiting order).  Binned-SAH top-down build, one shape per leaf, children
// boxes stored in the parent (2 box tests per node visit); large scenes are built by several threads, with a result
// that does not depend on how many.
#pragma once

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <future>
#include <limits>
#include <thread>
#include <vector>

namespace rt {

struct Aabb {
    double lo[3], hi[3];
    void reset() {
        for (int k = 0; k < 3; ++k) {
            lo[k] = std::numeric_limits<double>::infinity();
            hi[k] = -std::numeric_limits<double>::infinity();
        }
    }
    void grow(const Aabb& b) {
        for (int k = 0; k < 3; ++k) {
            lo[k] = std::min(lo[k], b.lo[k]);
            hi[k] = std::max(hi[k], b.hi[k]);
        }
    }
    void grow_point(const double p[3]) {
        for (int k = 0; k < 3; ++k) {
            lo[k] = std::min(lo[k], p[k]);
            hi[k] = std::max(hi[k], p[k]);
        }
    }
    double half_area() const {
        const double dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        if (!(dx >= 0 && dy >= 0 && dz >= 0)) return 0.0;
        return dx * dy + dy * dz + dz * dx;
    }
};

struct BvhNode {
    Aabb box[2];
    int32_t child[2];  // >= 0 inner node, < 0 leaf ~item (item = index into the builder's input order)
};

struct Bvh {
    std::vector<BvhNode> nodes;
    std::vector<uint32_t> leaf_order;  // items in depth-first leaf order
    int32_t root = 0;                  // node index, or ~0 for a single item
    int max_depth = 0;
};

// items: boxes of the bounded shapes.  Leaves reference positions in `leaf_order`.
//
// The tree is binary with one item per leaf, emitted in depth-first order, left subtree first: the subtree over the
// items [begin, end) of the (partitioned) index array owns exactly the nodes [r, r + count - 1) and the leaves
// [begin, end), where r is its root's index, its left child sits at r + 1 and its right child at r + (mid - begin).
// Every index is therefore known before anything below it is built, the arrays can be sized up front, and disjoint
// subtrees can be built by different threads into their own slices: the result does not depend on the thread count
// (tests/test_bvh_build.py).  threads = 0: as many as the host offers.
namespace bvh_detail {

struct Build {
    const std::vector<Aabb>& boxes;
    std::vector<uint32_t>& idx;
    const double* cen[3];
    Bvh& out;
    int depth_limit;
    std::atomic<int> max_depth{0};
    Build(const std::vector<Aabb>& b, std::vector<uint32_t>& i, Bvh& o, int limit) : boxes(b), idx(i), out(o), depth_limit(limit) {}
};

struct Task {
    uint32_t begin, end;
    uint32_t node;  // index of the subtree's root node (unused for a single item)
    int depth;
};

// Splits [begin, end) (count >= 2) and writes the node; returns the split position.
inline uint32_t split_range(Build& B, const Task& t) {
    std::vector<uint32_t>& idx = B.idx;
    const std::vector<Aabb>& boxes = B.boxes;
    const uint32_t count = t.end - t.begin;
    constexpr int BINS = 16;
    // centroid bounds -> split axis
    double clo[3] = {1e300, 1e300, 1e300}, chi[3] = {-1e300, -1e300, -1e300};
    for (uint32_t i = t.begin; i < t.end; ++i)
        for (int k = 0; k < 3; ++k) {
            clo[k] = std::min(clo[k], B.cen[k][idx[i]]);
            chi[k] = std::max(chi[k], B.cen[k][idx[i]]);
        }
    int axis = 0;
    if (chi[1] - clo[1] > chi[axis] - clo[axis]) axis = 1;
    if (chi[2] - clo[2] > chi[axis] - clo[axis]) axis = 2;
    uint32_t mid = t.begin + count / 2;
    const double extent = chi[axis] - clo[axis];
    bool split_done = false;
    if (extent > 0 && count > 4 && t.depth + (int)std::ceil(std::log2((double)count)) + 2 < B.depth_limit) {
        // binned SAH
        Aabb bb[BINS];
        uint32_t bc[BINS] = {0};
        for (auto& b : bb) b.reset();
        const double scale = BINS * (1.0 - 1e-9) / extent;
        for (uint32_t i = t.begin; i < t.end; ++i) {
            int b = (int)((B.cen[axis][idx[i]] - clo[axis]) * scale);
            b = std::min(std::max(b, 0), BINS - 1);
            bb[b].grow(boxes[idx[i]]);
            bc[b]++;
        }
        double right_area[BINS];
        uint32_t right_count[BINS];
        Aabb acc;
        acc.reset();
        uint32_t cnt = 0;
        for (int b = BINS - 1; b >= 1; --b) {
            acc.grow(bb[b]);
            cnt += bc[b];
            right_area[b] = acc.half_area();
            right_count[b] = cnt;
        }
        acc.reset();
        cnt = 0;
        double best = std::numeric_limits<double>::infinity();
        int best_split = -1;
        for (int b = 0; b < BINS - 1; ++b) {
            acc.grow(bb[b]);
            cnt += bc[b];
            if (cnt == 0 || right_count[b + 1] == 0) continue;
            const double cost = acc.half_area() * cnt + right_area[b + 1] * right_count[b + 1];
            if (cost < best) {
                best = cost;
                best_split = b;
            }
        }
        if (best_split >= 0) {
            const double thr_bin = best_split + 1;
            auto it = std::partition(idx.begin() + t.begin, idx.begin() + t.end, [&](uint32_t v) {
                int b = (int)((B.cen[axis][v] - clo[axis]) * scale);
                b = std::min(std::max(b, 0), BINS - 1);
                return b < thr_bin;
            });
            mid = (uint32_t)(it - idx.begin());
            split_done = mid > t.begin && mid < t.end;
        }
    }
    if (!split_done) {
        mid = t.begin + count / 2;
        std::nth_element(idx.begin() + t.begin, idx.begin() + mid, idx.begin() + t.end,
                         [&](uint32_t a, uint32_t b) { return B.cen[axis][a] < B.cen[axis][b]; });
    }
    BvhNode& node = B.out.nodes[t.node];
    node.box[0].reset();
    node.box[1].reset();
    for (uint32_t i = t.begin; i < mid; ++i) node.box[0].grow(boxes[idx[i]]);
    for (uint32_t i = mid; i < t.end; ++i) node.box[1].grow(boxes[idx[i]]);
    // children by the layout rule above: a single item is the leaf at its own position in the index array
    node.child[0] = mid - t.begin == 1 ? ~(int32_t)t.begin : (int32_t)(t.node + 1u);
    node.child[1] = t.end - mid == 1 ? ~(int32_t)mid : (int32_t)(t.node + (mid - t.begin));
    return mid;
}

// Builds the subtree of one task on the calling thread; subtrees of at least `spawn_min` items found on the way down are
// handed to other threads while fewer than `spawn_depth` splits lie above them.
inline void build_subtree(Build& B, Task root, uint32_t spawn_min, int spawn_depth) {
    std::vector<Task> stack;
    std::vector<std::future<void>> spawned;
    stack.push_back(root);
    int deepest = 0;
    while (!stack.empty()) {
        const Task t = stack.back();
        stack.pop_back();
        deepest = std::max(deepest, t.depth);
        if (t.end - t.begin == 1) {
            B.out.leaf_order[t.begin] = B.idx[t.begin];
            continue;
        }
        const uint32_t mid = split_range(B, t);
        const Task left{t.begin, mid, t.node + 1u, t.depth + 1}, right{mid, t.end, t.node + (mid - t.begin), t.depth + 1};
        if (t.depth - root.depth < spawn_depth && mid - t.begin >= spawn_min && t.end - mid >= spawn_min) {
            spawned.push_back(std::async(std::launch::async, [&B, left, spawn_min, spawn_depth, &root, t] {
                build_subtree(B, left, spawn_min, spawn_depth - (t.depth + 1 - root.depth));
            }));
            stack.push_back(right);
        } else {
            stack.push_back(right);
            stack.push_back(left);
        }
    }
    int seen = B.max_depth.load();
    while (deepest > seen && !B.max_depth.compare_exchange_weak(seen, deepest)) {
    }
    for (auto& f : spawned) f.get();
}

}  // namespace bvh_detail

inline Bvh build_bvh(const std::vector<Aabb>& boxes, int depth_limit, unsigned threads = 0) {
    Bvh bvh;
    const uint32_t n = (uint32_t)boxes.size();
    if (n == 0) return bvh;
    std::vector<uint32_t> idx(n);
    for (uint32_t i = 0; i < n; ++i) idx[i] = i;
    std::vector<double> cx(n), cy(n), cz(n);
    for (uint32_t i = 0; i < n; ++i) {
        cx[i] = 0.5 * (boxes[i].lo[0] + boxes[i].hi[0]);
        cy[i] = 0.5 * (boxes[i].lo[1] + boxes[i].hi[1]);
        cz[i] = 0.5 * (boxes[i].lo[2] + boxes[i].hi[2]);
    }
    bvh.leaf_order.assign(n, 0u);
    if (n == 1) {
        bvh.root = ~0;
        return bvh;
    }
    bvh.nodes.resize(n - 1);
    bvh.root = 0;
    bvh_detail::Build B(boxes, idx, bvh, depth_limit);
    B.cen[0] = cx.data();
    B.cen[1] = cy.data();
    B.cen[2] = cz.data();
    if (threads == 0) threads = std::max(1u, std::thread::hardware_concurrency());
    // up to 2^spawn_depth subtrees in flight; small scenes are built on the calling thread
    int spawn_depth = 0;
    while ((1u << spawn_depth) < 2u * threads && spawn_depth < 8) ++spawn_depth;
    if (threads <= 1 || n < (1u << 15)) spawn_depth = 0;
    bvh_detail::build_subtree(B, bvh_detail::Task{0u, n, 0u, 1}, std::max(1u << 12, n >> (spawn_depth + 2)), spawn_depth);
    bvh.max_depth = B.max_depth.load();
    return bvh;
}

}  // namespace rt
