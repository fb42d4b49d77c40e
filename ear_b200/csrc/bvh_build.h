// Host-side BVH construction for the GPU traversal kernels (K0 in SURVEY.md section 2.3).
// The reference has no acceleration structure (Mesh::RayIntersection loops over every triangle,
// src/Mesh.cpp:39); this replaces that O(T) loop while returning the SAME winner.
#pragma once
#include <cstdint>
#include <memory>
#include <utility>
#include <vector>

namespace earb {

// One 4-wide node (64 bytes = two 256-bit loads).  The SAH tree is built binary and collapsed (largest-area
// child expanded first); each child's padded box is quantised to 8 bits per plane inside the node's own
// bounds, rounded OUTWARD, so a decoded box always contains the float box it came from:
//   plane(q) = lo[a] + q * 2^(ex[a]-127)
//   lo[3]      node bounds, low corner (float32)
//   ex[3]      biased float exponents of the per-axis grid step (255 steps cover the extent)
//   child[4]   >= 0: node index; < 0: leaf, ~child = first*8 + (count-1); kEmptyChild: nothing there
//   q[p][k]    plane p (lo.x, lo.y, lo.z, hi.x, hi.y, hi.z) of child k; empty children are inverted (255 / 0)
//   slack[4]   fp16, rounded up: per-child ray-interval widening of the EXACT mode (DESIGN.md "exactness")
struct Node {
	float lo[3];
	uint8_t ex[3], pad;
	int32_t child[4];
	uint8_t q[6][4];
	uint16_t slack[4];
};
static_assert(sizeof(Node) == 64, "node must be 64 bytes");

// Triangle record in leaf order (64 bytes):
//   (v0.xyz, original index) (e1.xyz, material) (e2.xyz, 0) (unit normal, 0)
// e1 = v1 - v0, e2 = v2 - v0 in float32, exactly the edge vectors gmtl::intersectDoubleSided forms
// per test; normal = normalize(e1 x e2) in the reference's float32 order (Triangle ctor,
// src/Triangle.cpp:41) -- this file must be compiled without FMA contraction.
struct TriRecord {
	float v0[3]; int32_t index;
	float e1[3]; int32_t material;
	float e2[3]; int32_t pad0;
	float normal[3]; int32_t pad1;
};
static_assert(sizeof(TriRecord) == 64, "triangle record must be 64 bytes");

constexpr int32_t kEmptyChild = 0x7fffffff;
constexpr int kMaxLeaf = 4;

// Allocator whose resize() leaves trivially-constructible elements uninitialised: the builder overwrites every
// element, and zero-filling 100+ MB first costs more than some of its passes.
template <class T>
struct RawAllocator {
	using value_type = T;
	RawAllocator() = default;
	template <class U> RawAllocator(const RawAllocator<U>&) {}
	T* allocate(std::size_t n) { return std::allocator<T>().allocate(n); }
	void deallocate(T* p, std::size_t n) { std::allocator<T>().deallocate(p, n); }
	template <class U> void construct(U* p) { ::new (static_cast<void*>(p)) U; }
	template <class U, class A0, class... A> void construct(U* p, A0&& a0, A&&... a) { ::new (static_cast<void*>(p)) U(std::forward<A0>(a0), std::forward<A>(a)...); }
	template <class U> bool operator==(const RawAllocator<U>&) const { return true; }
	template <class U> bool operator!=(const RawAllocator<U>&) const { return false; }
};
template <class T> using RawVector = std::vector<T, RawAllocator<T>>;

struct Bvh {
	RawVector<Node> nodes;           // nodes[0] is the root and is always internal
	RawVector<TriRecord> tris;       // leaf order
	float lo[3], hi[3];              // scene bounds (unpadded)
	float diagonal;
	float s0;                        // absolute ray-interval margin used with the relative slack
	int depth;
};

// verts: [n][3][3] float32 in file order.  Deterministic for a given input.
void build_bvh(const float* verts, const int32_t* tri_material, int32_t n, Bvh& out);

}  // namespace earb
