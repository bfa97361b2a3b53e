#include "simt_shim.h"
extern "C" void K2
(float *Vedge_weight, float *Vhinb, float *Vnormcen, float *Vnorminb, float *V11, 
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
            
            float V10_tmp = 0;
            int offset3 = dst_id * 1 + tx/7;int offset4 = dst_id * 7 + tx;
            
            for (int e=beg;e<end;++e) {
                
                int src_id = __ldg(column_indices + e);
                int eid = __ldg(eids + e);
                
                int offset0 = src_id * 1 + tx/7;int offset1 = src_id * 7 + tx;int offset2 = eid * 1 + tx/7;
                
                
                
                float V8_tmp = Vnorminb[offset0]*Vhinb[offset1];
                
                
                
                float V9_tmp = V8_tmp*Vedge_weight[offset2];
                
                
                
                
                V10_tmp += V9_tmp;
                
                
            }
            
            
            
            
            
            
            
            float V11_tmp = V10_tmp*Vnormcen[offset3];
            V11[offset4] = V11_tmp;
            
        }
    }
}extern "C" void K3
(float *V12, float *Vedge_weight, float *Vnormcen, float *Vnorminb, float *V17, 
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
            
            float V16_tmp = 0;
            int offset3 = src_id * 1 + tx/7;int offset4 = src_id * 7 + tx;
            
            for (int e=beg;e<end;++e) {
                
                int dst_id = __ldg(column_indices + e);
                int eid = __ldg(eids + e);
                
                int offset0 = dst_id * 7 + tx;int offset1 = dst_id * 1 + tx/7;int offset2 = eid * 1 + tx/7;
                
                
                
                float V13_tmp = V12[offset0]*Vnormcen[offset1];
                
                
                
                float V15_tmp = V13_tmp*Vedge_weight[offset2];
                
                
                
                
                V16_tmp += V15_tmp;
                
                
            }
            
            
            
            
            
            
            
            float V17_tmp = V16_tmp*Vnorminb[offset3];
            V17[offset4] = V17_tmp;
            
        }
    }
}

extern "C" void run_K2(void** t, int* row_offsets, int* eids, int* column_indices, int* node_ids,
    int num_nodes, int max_dimx, int max_dimy, int thrs_per_group, int nodes_per_block, int nblks, int nthrs) {
  blockDim.x = nthrs; blockDim.y = 1; blockDim.z = 1; gridDim.x = nblks;
  for (int b = 0; b < nblks; ++b) for (int th = 0; th < nthrs; ++th) {
    blockIdx.x = b; threadIdx.x = th;
    K2((float*)t[0], (float*)t[1], (float*)t[2], (float*)t[3], (float*)t[4], row_offsets, eids, column_indices, node_ids, num_nodes, max_dimx, max_dimy, thrs_per_group, nodes_per_block);
  }
}
extern "C" void run_K3(void** t, int* row_offsets, int* eids, int* column_indices, int* node_ids,
    int num_nodes, int max_dimx, int max_dimy, int thrs_per_group, int nodes_per_block, int nblks, int nthrs) {
  blockDim.x = nthrs; blockDim.y = 1; blockDim.z = 1; gridDim.x = nblks;
  for (int b = 0; b < nblks; ++b) for (int th = 0; th < nthrs; ++th) {
    blockIdx.x = b; threadIdx.x = th;
    K3((float*)t[0], (float*)t[1], (float*)t[2], (float*)t[3], (float*)t[4], row_offsets, eids, column_indices, node_ids, num_nodes, max_dimx, max_dimy, thrs_per_group, nodes_per_block);
  }
}