// Test helper: the host-callable half of camus_b200/csrc/layout.cuh behind a C ABI, compiled with
// g++ by tests/test_layout_mirror.py so that camus_b200/partition.py can be checked against the
// very functions the CUDA library uses (no GPU involved).
#include "../camus_b200/csrc/layout.cuh"

using namespace camus_dev;

extern "C" {
unsigned probe_seg_class(unsigned s) { return seg_class(s); }
unsigned probe_class_threshold(unsigned m, int world) { return class_threshold(m, world); }
unsigned probe_class_need_mask(int rank, int world) { return class_need_mask(rank, world); }
int probe_class_owner(unsigned sa, unsigned sb, unsigned sc, unsigned sd, unsigned m, int world) {
  return class_owner(seg_class(sa), seg_class(sb), class_hash(sc, sd), world, class_threshold(m, world));
}
unsigned long long probe_box_index(unsigned sa, unsigned sb, unsigned sc, unsigned sd) { return box_index(sa, sb, sc, sd); }
void probe_box_unrank(unsigned long long r, unsigned* out) { box_unrank(r, out[0], out[1], out[2], out[3]); }
unsigned long long probe_tile_index(unsigned s0, unsigned s1, unsigned s2) { return tile_index(s0, s1, s2); }
unsigned long long probe_partition_chunk(unsigned long long nb, unsigned world) { return partition_chunk(nb, world); }
unsigned long long probe_local_box_count(unsigned long long nb, unsigned rank, unsigned world) {
  return local_box_count(nb, partition_chunk(nb, world), rank, world);
}
// plane work list: returns the number of entries; out (may be null) receives s0, s1, s2, slot per entry
unsigned long long probe_plane_work_list(unsigned m, unsigned need, int compact, unsigned* out) {
  const std::vector<PlaneDesc> l = plane_work_list(m, need, compact != 0);
  if (out)
    for (size_t i = 0; i < l.size(); ++i) {
      out[4 * i + 0] = l[i].s01 & 0xFFFFu;
      out[4 * i + 1] = l[i].s01 >> 16;
      out[4 * i + 2] = l[i].s2;
      out[4 * i + 3] = l[i].slot;
    }
  return l.size();
}
int probe_slice_off(int w, int r, int q) { return w == 8 ? slice_off<8>(r, q) : slice_off<12>(r, q); }
int probe_slice_array_words(int w) { return w == 8 ? SliceArray<8>::words : SliceArray<12>::words; }
}
