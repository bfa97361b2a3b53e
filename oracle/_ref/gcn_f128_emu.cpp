#include "simt_shim.h"
extern "C" void K14
(float *Vhinb, float *Vnormcen, float *Vnorminb, float *V104, 
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
            
            float V103_tmp = 0;
            int offset2 = dst_id * 1 + tx/128;int offset3 = dst_id * 128 + tx;
            
            for (int e=beg;e<end;++e) {
                
                int src_id = __ldg(column_indices + e);
                int eid = __ldg(eids + e);
                
                int offset0 = src_id * 1 + tx/128;int offset1 = src_id * 128 + tx;
                
                
                
                float V102_tmp = Vhinb[offset1]*Vnorminb[offset0];
                
                
                
                
                V103_tmp += V102_tmp;
                
                
            }
            
            
            
            
            
            
            
            float V104_tmp = V103_tmp*Vnormcen[offset2];
            V104[offset3] = V104_tmp;
            
        }
    }
}extern "C" void K15
(float *V105, float *Vnormcen, float *Vnorminb, float *V109, 
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
            
            float V108_tmp = 0;
            int offset2 = src_id * 1 + tx/128;int offset3 = src_id * 128 + tx;
            
            for (int e=beg;e<end;++e) {
                
                int dst_id = __ldg(column_indices + e);
                int eid = __ldg(eids + e);
                
                int offset0 = dst_id * 1 + tx/128;int offset1 = dst_id * 128 + tx;
                
                
                
                float V106_tmp = V105[offset1]*Vnormcen[offset0];
                
                
                
                
                V108_tmp += V106_tmp;
                
                
            }
            
            
            
            
            
            
            
            float V109_tmp = V108_tmp*Vnorminb[offset2];
            V109[offset3] = V109_tmp;
            
        }
    }
}

extern "C" void run_K14(void** t, int* row_offsets, int* eids, int* column_indices, int* node_ids,
    int num_nodes, int max_dimx, int max_dimy, int thrs_per_group, int nodes_per_block, int nblks, int nthrs) {
  blockDim.x = nthrs; blockDim.y = 1; blockDim.z = 1; gridDim.x = nblks;
  for (int b = 0; b < nblks; ++b) for (int th = 0; th < nthrs; ++th) {
    blockIdx.x = b; threadIdx.x = th;
    K14((float*)t[0], (float*)t[1], (float*)t[2], (float*)t[3], row_offsets, eids, column_indices, node_ids, num_nodes, max_dimx, max_dimy, thrs_per_group, nodes_per_block);
  }
}
extern "C" void run_K15(void** t, int* row_offsets, int* eids, int* column_indices, int* node_ids,
    int num_nodes, int max_dimx, int max_dimy, int thrs_per_group, int nodes_per_block, int nblks, int nthrs) {
  blockDim.x = nthrs; blockDim.y = 1; blockDim.z = 1; gridDim.x = nblks;
  for (int b = 0; b < nblks; ++b) for (int th = 0; th < nthrs; ++th) {
    blockIdx.x = b; threadIdx.x = th;
    K15((float*)t[0], (float*)t[1], (float*)t[2], (float*)t[3], row_offsets, eids, column_indices, node_ids, num_nodes, max_dimx, max_dimy, thrs_per_group, nodes_per_block);
  }
}