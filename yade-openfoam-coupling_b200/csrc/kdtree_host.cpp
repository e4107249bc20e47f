// kdtree_host.cpp -- one-time host construction of the k-d tree over the cell centres.
//
// Follows the construction RULE of the reference (FoamYade/meshtree/meshTree.C:9-51):
//   * points enter in cell-index order (MT.C:13-14),
//   * at depth d the split axis is d % 3 (MT.C:24),
//   * the node is the element std::nth_element leaves at index size/2 under a strict `<` on that
//     axis' coordinate (MT.C:46-51, comparator MT.H:45-55),
//   * left subtree = the elements before it, right subtree = the elements after it, each in the
//     order nth_element left them (MT.C:30-31).
// Which of many coordinate-tied lattice points becomes the median is decided by libstdc++'s
// introselect, so the cell lists the traversal returns are only reproducible if the same library
// routine is driven by the same comparison sequence; hence std::nth_element here, on compact
// 32-byte records and in place (the reference copies both halves at every level and heap-allocates
// three pointers per point, which is where its 6-7 s at 128^3 go).
//
// The permuted array IS the tree: the node of the range [lo,hi) is element lo + (hi-lo)/2.
#include <algorithm>
#include <utility>
#include <vector>

#include "fy_ctx.h"

namespace {
struct AxisLess {
    int a;
    bool operator()(const FyKdNode& p, const FyKdNode& q) const
    {
        const double pv = a == 0 ? p.x : (a == 1 ? p.y : p.z);
        const double qv = a == 0 ? q.x : (a == 1 ? q.y : q.z);
        return pv < qv;
    }
};
}

void fyBuildKdTree(const double* C, int n, std::vector<FyKdNode>& t)
{
    t.resize((size_t)n);
    for (int i = 0; i < n; ++i) {
        t[i].x = C[3 * (size_t)i];
        t[i].y = C[3 * (size_t)i + 1];
        t[i].z = C[3 * (size_t)i + 2];
        t[i].id = i;
        t[i].pad = 0;
    }
    // explicit stack instead of recursion: (lo, hi, depth)
    struct Job { int lo, hi, depth; };
    std::vector<Job> st;
    st.push_back(Job{0, n, 0});
    while (!st.empty()) {
        const Job j = st.back();
        st.pop_back();
        const int sz = j.hi - j.lo;
        if (sz <= 1) continue;      // nth_element on one element is a no-op
        const int md = j.lo + sz / 2;
        std::nth_element(t.begin() + j.lo, t.begin() + md, t.begin() + j.hi, AxisLess{j.depth % 3});
        st.push_back(Job{md + 1, j.hi, j.depth + 1});
        st.push_back(Job{j.lo, md, j.depth + 1});
    }
}
