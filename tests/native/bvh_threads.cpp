// Harness for tests/test_bvh_build.py: the BVH builder (csrc/rt_bvh.h) must give the same tree for any thread count,
// and a valid one: every item in exactly one leaf, node boxes enclosing their subtrees, the documented preorder layout.
#include <chrono>
#include <cstdio>
#include <cstring>
#include <functional>
#include <random>

#include "rt_bvh.h"

static bool encloses(const rt::Aabb& outer, const rt::Aabb& inner) {
    for (int k = 0; k < 3; ++k)
        if (!(outer.lo[k] <= inner.lo[k] && outer.hi[k] >= inner.hi[k])) return false;
    return true;
}

int main(int argc, char** argv) {
    const unsigned n = argc > 1 ? (unsigned)atoi(argv[1]) : 100000u;
    const int kind = argc > 2 ? atoi(argv[2]) : 0;  // 0 uniform, 1 clustered with many identical centroids
    std::mt19937_64 rng(1234 + n);
    std::uniform_real_distribution<double> u(-100, 100), r(0.01, 2.0);
    std::vector<rt::Aabb> boxes(n);
    for (unsigned i = 0; i < n; ++i) {
        double c[3] = {u(rng), u(rng) * 0.2, u(rng)}, h = r(rng);
        if (kind == 1) {
            for (int k = 0; k < 3; ++k) c[k] = std::floor(c[k] / 25.0) * 25.0;  // a 9 x 3 x 9 grid of coincident centres
            h = 1.0;
        }
        for (int k = 0; k < 3; ++k) {
            boxes[i].lo[k] = c[k] - h;
            boxes[i].hi[k] = c[k] + h;
        }
    }
    auto t0 = std::chrono::steady_clock::now();
    const rt::Bvh one = rt::build_bvh(boxes, 60, 1);
    auto t1 = std::chrono::steady_clock::now();
    const rt::Bvh many = rt::build_bvh(boxes, 60, 0), three = rt::build_bvh(boxes, 60, 3);
    auto t2 = std::chrono::steady_clock::now();
    auto equal = [](const rt::Bvh& a, const rt::Bvh& b) {
        return a.root == b.root && a.max_depth == b.max_depth && a.leaf_order == b.leaf_order && a.nodes.size() == b.nodes.size() &&
               (a.nodes.empty() || memcmp(a.nodes.data(), b.nodes.data(), a.nodes.size() * sizeof(rt::BvhNode)) == 0);
    };
    bool ok = equal(one, many) && equal(one, three);
    // validity: a permutation in the leaves, boxes enclose, preorder layout (left child at r + 1, subtree of c items = c - 1 nodes)
    std::vector<unsigned char> seen(n, 0);
    for (uint32_t item : one.leaf_order) {
        if (item >= n || seen[item]) ok = false;
        else seen[item] = 1;
    }
    if (one.leaf_order.size() != n || one.nodes.size() != (n ? n - 1 : 0)) ok = false;
    std::function<uint32_t(int32_t, rt::Aabb*)> walk = [&](int32_t ref, rt::Aabb* box) -> uint32_t {  // returns the item count below ref
        if (ref < 0) {
            *box = boxes[one.leaf_order[(uint32_t)~ref]];
            return 1u;
        }
        const rt::BvhNode& nd = one.nodes[(uint32_t)ref];
        rt::Aabb l, r2;
        const uint32_t cl = walk(nd.child[0], &l), cr = walk(nd.child[1], &r2);
        if (!encloses(nd.box[0], l) || !encloses(nd.box[1], r2)) ok = false;
        if (nd.child[0] >= 0 && nd.child[0] != ref + 1) ok = false;
        if (nd.child[1] >= 0 && (uint32_t)nd.child[1] != (uint32_t)ref + cl) ok = false;
        *box = nd.box[0];
        box->grow(nd.box[1]);
        return cl + cr;
    };
    if (n >= 2) {
        rt::Aabb all;
        if (walk(one.root, &all) != n) ok = false;
    }
    auto ms = [](auto x, auto y) { return std::chrono::duration<double, std::milli>(y - x).count(); };
    printf("n %u kind %d: 1 thread %.1f ms, %u + 3 threads %.1f ms, depth %d, %s\n", n, kind, ms(t0, t1), std::thread::hardware_concurrency(), ms(t1, t2),
           one.max_depth, ok ? "ok" : "MISMATCH");
    return ok ? 0 : 1;
}
