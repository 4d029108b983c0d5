// Sampler / to_block host entry points shared with plan.cu.
#pragma once
#include "graph.cuh"

namespace ogl {

struct ToBlockWs {
  int32_t* first = nullptr;         // [v_cap] first-appearance table, 0x7fffffff when idle
  int32_t* flags = nullptr;         // [ne_max]
  int32_t* pos = nullptr;           // [ne_max]
  int32_t* scan_scratch = nullptr;
  int32_t* n_new = nullptr;
  int32_t* rev_cnt = nullptr;       // [2 * rev_cap]: per-source counters, then fill cursors
  int32_t* rev_scan_scratch = nullptr;
  int32_t* rev_sort_scratch = nullptr;   // [ne_max + 2]: [0] = number of long rows, [1 ..] their ids, then a copy buffer
  int64_t rev_cap = 0;
  int64_t v_cap = 0, ne_max = 0;
};

int to_block_init(ToBlockWs* ws, int64_t v_cap, int64_t ne_max, int64_t rows_max);
void to_block_free(ToBlockWs* ws);

// picks for rows dst_nodes[0 .. *n_dst_dev) (n_dst_dev == nullptr: n_dst_max rows); step from step_dev if non-null
int sample_hop(const GraphView& g, const int32_t* dst_nodes, const int32_t* n_dst_dev, int n_dst_max, int fanout, uint64_t seed,
               const uint32_t* step_dev, uint32_t step_imm, uint32_t hop, int32_t* out_src, int64_t* out_eid, cudaStream_t s);

// src_nodes = dst_nodes ++ first-appearance-ordered new sources; edge_lid[p] = local id of picked[p] (-1 if empty)
int to_block(ToBlockWs* ws, const int32_t* dst_nodes, const int32_t* n_dst_dev, int n_dst_max, int fanout, const int32_t* picked,
             int32_t* src_nodes, int32_t* n_src_dev, int n_src_max, int32_t* edge_lid, cudaStream_t s);

// reverse edge lists of a sampled block: rev_ptr[n_src_max + 1] (exclusive offsets), rev_edge[k] = (destination row d << 8) | slot j   (fan-outs < 255, d < 2^23)
int reverse_edges(ToBlockWs* ws, const int32_t* edge_lid, const int32_t* n_dst_dev, int n_dst_max, int fanout, int n_src_max,
                  int32_t* rev_ptr, int32_t* rev_edge, cudaStream_t s);

int cast_nodes(const int64_t* in, int32_t* out, int64_t n, int64_t n_vertices, uint32_t* err_flag, cudaStream_t s);

}  // namespace ogl
