// Host-side BVH construction for the GPU traversal kernels (K0 in SURVEY.md section 2.3).
// The reference has no acceleration structure (Mesh::RayIntersection loops over every triangle,
// src/Mesh.cpp:39); this replaces that O(T) loop while returning the SAME winner.
#pragma once
#include <cstdint>
#include <vector>

namespace earb {

// One binary node, both children's boxes inline (64 bytes, four float4 loads):
//   a = (c0.lo.x, c0.hi.x, c0.lo.y, c0.hi.y)
//   b = (c1.lo.x, c1.hi.x, c1.lo.y, c1.hi.y)
//   c = (c0.lo.z, c0.hi.z, c1.lo.z, c1.hi.z)
//   d = (child0, child1, slack0, slack1)   child >= 0: node index; child < 0: leaf, ~child = first*8 + (count-1)
//                                          kEmptyChild: nothing there
// slackN (float bits) widens the ray interval used to cull that child; see DESIGN.md "exactness".
struct Node {
	float a[4], b[4], c[4];
	int32_t child[2];
	float slack[2];
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

struct Bvh {
	std::vector<Node> nodes;         // nodes[0] is the root and is always internal
	std::vector<TriRecord> tris;     // leaf order
	float lo[3], hi[3];              // scene bounds (unpadded)
	float diagonal;
	float s0;                        // absolute ray-interval margin used with the relative slack
	int depth;
};

// verts: [n][3][3] float32 in file order.  Deterministic for a given input.
void build_bvh(const float* verts, const int32_t* tri_material, int32_t n, Bvh& out);

}  // namespace earb
