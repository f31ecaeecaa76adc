// rt_bvh.h — host-side BVH builder over the world-space boxes of the bounded shapes.
//
// Not part of the reference (it tests every shape for every ray, composites/world.rs:31-33).  The
// hierarchy only decides WHICH exact ray-shape tests run; every test that can produce an
// intersection the query cares about still runs, in the reference's arithmetic, so results are
// unchanged (boxes are inflated far beyond f64 rounding, and hit selection is by (distance, world
// order), independent of visiting order).  Binned-SAH top-down build, one shape per leaf, children
// boxes stored in the parent (2 box tests per node visit).
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
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
inline Bvh build_bvh(const std::vector<Aabb>& boxes, int depth_limit) {
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
    const double* cen[3] = {cx.data(), cy.data(), cz.data()};
    bvh.leaf_order.reserve(n);
    if (n == 1) {
        bvh.leaf_order.push_back(0);
        bvh.root = ~0;
        return bvh;
    }
    bvh.nodes.reserve(n - 1);

    struct Task {
        uint32_t begin, end;
        int32_t parent;  // node whose child slot receives the result, -1 for the root
        int slot;
        int depth;
    };
    // depth-first with an explicit stack; right child pushed first so the left subtree is emitted first
    std::vector<Task> stack;
    stack.push_back({0, n, -1, 0, 1});
    constexpr int BINS = 16;
    while (!stack.empty()) {
        Task t = stack.back();
        stack.pop_back();
        const uint32_t count = t.end - t.begin;
        bvh.max_depth = std::max(bvh.max_depth, t.depth);
        int32_t ref;
        if (count == 1) {
            ref = ~(int32_t)bvh.leaf_order.size();
            bvh.leaf_order.push_back(idx[t.begin]);
        } else {
            // centroid bounds -> split axis
            double clo[3] = {1e300, 1e300, 1e300}, chi[3] = {-1e300, -1e300, -1e300};
            for (uint32_t i = t.begin; i < t.end; ++i)
                for (int k = 0; k < 3; ++k) {
                    clo[k] = std::min(clo[k], cen[k][idx[i]]);
                    chi[k] = std::max(chi[k], cen[k][idx[i]]);
                }
            int axis = 0;
            if (chi[1] - clo[1] > chi[axis] - clo[axis]) axis = 1;
            if (chi[2] - clo[2] > chi[axis] - clo[axis]) axis = 2;
            uint32_t mid = t.begin + count / 2;
            const double extent = chi[axis] - clo[axis];
            bool split_done = false;
            if (extent > 0 && count > 4 && t.depth + (int)std::ceil(std::log2((double)count)) + 2 < depth_limit) {
                // binned SAH
                Aabb bb[BINS];
                uint32_t bc[BINS] = {0};
                for (auto& b : bb) b.reset();
                const double scale = BINS * (1.0 - 1e-9) / extent;
                for (uint32_t i = t.begin; i < t.end; ++i) {
                    int b = (int)((cen[axis][idx[i]] - clo[axis]) * scale);
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
                        int b = (int)((cen[axis][v] - clo[axis]) * scale);
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
                                 [&](uint32_t a, uint32_t b) { return cen[axis][a] < cen[axis][b]; });
            }
            ref = (int32_t)bvh.nodes.size();
            BvhNode node;
            node.box[0].reset();
            node.box[1].reset();
            for (uint32_t i = t.begin; i < mid; ++i) node.box[0].grow(boxes[idx[i]]);
            for (uint32_t i = mid; i < t.end; ++i) node.box[1].grow(boxes[idx[i]]);
            node.child[0] = node.child[1] = 0;
            bvh.nodes.push_back(node);
            stack.push_back({mid, t.end, ref, 1, t.depth + 1});
            stack.push_back({t.begin, mid, ref, 0, t.depth + 1});
        }
        if (t.parent < 0) bvh.root = ref;
        else bvh.nodes[t.parent].child[t.slot] = ref;
    }
    return bvh;
}

}  // namespace rt
