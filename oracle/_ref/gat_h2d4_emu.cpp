#include "simt_shim.h"
extern "C" void K8
(float *Velinb, float *Vercen, float *V59, float *V60, 
  int *row_offsets,
  int *eids,
  int *column_indices,
  int *node_ids,
  int num_nodes,
  int max_dimx,
  int max_dimy,
  int thrs_per_group,
  int nodes_per_block) {
      
    int dst_id = nodes_per_block*blockIdx.x + threadIdx.x/thrs_per_group;

    if (dst_id < num_nodes) {
        
        int feat_len = max_dimx * max_dimy;
        int beg = __ldg(row_offsets + dst_id);
        int end = __ldg(row_offsets + dst_id + 1);
        int tx = threadIdx.x % thrs_per_group;
        
        for (; tx<feat_len; tx+=blockDim.x) {
            
            float V60_tmp = 0;
            int offset0 = dst_id * 2 + tx;
            
            for (int e=beg;e<end;++e) {
                
                int src_id = __ldg(column_indices + e);
                int eid = __ldg(eids + e);
                
                int offset1 = src_id * 2 + tx;int offset2 = eid * 2 + tx;
                
                
                
                float V56_tmp = Velinb[offset1] + Vercen[offset0];
                
                
                
                float V57_tmp = V56_tmp - V56_tmp;
                
                
                
                float V58_tmp=V57_tmp>0?V57_tmp:0.2*V57_tmp;
                
                
                
                float V59_tmp = exp(V58_tmp);
                V59[offset2] = V59_tmp;
                
                
                
                V60_tmp += V59_tmp;
                
                
            }
            
            
            V60[offset0] = V60_tmp;
            
            
            
        }
    }
}extern "C" void K9
(float *V59, float *V60, float *Vfeat_srcinb, float *V63, 
  int *row_offsets,
  int *eids,
  int *column_indices,
  int *node_ids,
  int num_nodes,
  int max_dimx,
  int max_dimy,
  int thrs_per_group,
  int nodes_per_block) {
      
    int dst_id = nodes_per_block*blockIdx.x + threadIdx.x/thrs_per_group;

    if (dst_id < num_nodes) {
        
        int feat_len = max_dimx * max_dimy;
        int beg = __ldg(row_offsets + dst_id);
        int end = __ldg(row_offsets + dst_id + 1);
        int tx = threadIdx.x % thrs_per_group;
        
        for (; tx<feat_len; tx+=blockDim.x) {
            
            float V63_tmp = 0;
            int offset1 = dst_id * 2 + tx/4;int offset3 = dst_id * 8 + tx;
            
            for (int e=beg;e<end;++e) {
                
                int src_id = __ldg(column_indices + e);
                int eid = __ldg(eids + e);
                
                int offset2 = src_id * 8 + tx;int offset0 = eid * 2 + tx/4;
                
                
                
                float V61_tmp = V59[offset0]/V60[offset1];
                
                
                
                float V62_tmp = V61_tmp*Vfeat_srcinb[offset2];
                
                
                
                
                V63_tmp += V62_tmp;
                
                
            }
            
            
            V63[offset3] = V63_tmp;
            
            
            
        }
    }
}extern "C" void K10
(float *V59, float *V60, float *V63, float *V64, float *Velinb, float *Vercen, float *Vfeat_srcinb, float *V69, float *V85, float *V87, 
  int *row_offsets,
  int *eids,
  int *column_indices,
  int *node_ids,
  int num_nodes,
  int max_dimx,
  int max_dimy,
  int thrs_per_group,
  int nodes_per_block) {
      
    int src_id = nodes_per_block*blockIdx.x + threadIdx.x/thrs_per_group;

    if (src_id < num_nodes) {
        
        int feat_len = max_dimx * max_dimy;
        int beg = __ldg(row_offsets + src_id);
        int end = __ldg(row_offsets + src_id + 1);
        int tx = threadIdx.x % thrs_per_group;
        
        for (; tx<feat_len; tx+=blockDim.x) {
            
            float V85_tmp = 0;float V87_tmp = 0;float V69_tmp = 0;
            int offset3 = src_id * 2 + tx/4;int offset4 = src_id * 8 + tx;
            
            for (int e=beg;e<end;++e) {
                
                int dst_id = __ldg(column_indices + e);
                int eid = __ldg(eids + e);
                
                int offset1 = dst_id * 2 + tx/4;int offset2 = dst_id * 8 + tx;int offset0 = eid * 2 + tx/4;
                
                
                
                float V61_tmp = V59[offset0]/V60[offset1];
                
                
                
                float V68_tmp = V64[offset2]*V61_tmp;
                
                
                
                float V56_tmp = Velinb[offset3] + Vercen[offset1];
                
                
                
                float V57_tmp = V56_tmp - V56_tmp;
                
                
                
                float V66_tmp = V64[offset2]*Vfeat_srcinb[offset4];
                
                
                
                float V70_tmp = 1/V60[offset1];
                
                
                
                float V71_tmp = V66_tmp*V70_tmp;
                
                
                
                float V72_tmp = V64[offset2]/V60[offset1];
                
                
                
                float V73_tmp = V72_tmp*V63[offset2];
                
                
                
                float V74_tmp = -1*V73_tmp;
                
                
                
                float V78_tmp = V71_tmp + V74_tmp;
                
                
                
                float V79_tmp = V78_tmp*V59[offset0];
                
                
                
                float V80_tmp = V57_tmp>0?1:0.2;
                
                
                
                float V81_tmp = V79_tmp*V80_tmp;
                
                
                
                
                V85_tmp += V81_tmp;
                
                
                V87_tmp = V81_tmp;
                atomicAdd(V87+offset1, V87_tmp);
                
                V69_tmp += V68_tmp;
                
                
            }
            
            
            atomicAdd(V85+offset3, V85_tmp);
            
            
            
            V69[offset4] = V69_tmp;
            
            
            
        }
    }
}

extern "C" void run_K8(void** t, int* row_offsets, int* eids, int* column_indices, int* node_ids,
    int num_nodes, int max_dimx, int max_dimy, int thrs_per_group, int nodes_per_block, int nblks, int nthrs) {
  blockDim.x = nthrs; blockDim.y = 1; blockDim.z = 1; gridDim.x = nblks;
  for (int b = 0; b < nblks; ++b) for (int th = 0; th < nthrs; ++th) {
    blockIdx.x = b; threadIdx.x = th;
    K8((float*)t[0], (float*)t[1], (float*)t[2], (float*)t[3], row_offsets, eids, column_indices, node_ids, num_nodes, max_dimx, max_dimy, thrs_per_group, nodes_per_block);
  }
}
extern "C" void run_K9(void** t, int* row_offsets, int* eids, int* column_indices, int* node_ids,
    int num_nodes, int max_dimx, int max_dimy, int thrs_per_group, int nodes_per_block, int nblks, int nthrs) {
  blockDim.x = nthrs; blockDim.y = 1; blockDim.z = 1; gridDim.x = nblks;
  for (int b = 0; b < nblks; ++b) for (int th = 0; th < nthrs; ++th) {
    blockIdx.x = b; threadIdx.x = th;
    K9((float*)t[0], (float*)t[1], (float*)t[2], (float*)t[3], row_offsets, eids, column_indices, node_ids, num_nodes, max_dimx, max_dimy, thrs_per_group, nodes_per_block);
  }
}
extern "C" void run_K10(void** t, int* row_offsets, int* eids, int* column_indices, int* node_ids,
    int num_nodes, int max_dimx, int max_dimy, int thrs_per_group, int nodes_per_block, int nblks, int nthrs) {
  blockDim.x = nthrs; blockDim.y = 1; blockDim.z = 1; gridDim.x = nblks;
  for (int b = 0; b < nblks; ++b) for (int th = 0; th < nthrs; ++th) {
    blockIdx.x = b; threadIdx.x = th;
    K10((float*)t[0], (float*)t[1], (float*)t[2], (float*)t[3], (float*)t[4], (float*)t[5], (float*)t[6], (float*)t[7], (float*)t[8], (float*)t[9], row_offsets, eids, column_indices, node_ids, num_nodes, max_dimx, max_dimy, thrs_per_group, nodes_per_block);
  }
}