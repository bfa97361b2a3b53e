extern "C" __global__ void K4
(float *Velinb, float *Vercen, float *V21, float *V22, 
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
            
            float V22_tmp = 0;
            int offset0 = dst_id * 8 + tx;
            
            for (int e=beg;e<end;++e) {
                
                int src_id = __ldg(column_indices + e);
                int eid = __ldg(eids + e);
                
                int offset1 = src_id * 8 + tx;int offset2 = eid * 8 + tx;
                
                
                
                float V18_tmp = Velinb[offset1] + Vercen[offset0];
                
                
                
                float V19_tmp = V18_tmp - V18_tmp;
                
                
                
                float V20_tmp=V19_tmp>0?V19_tmp:0.2*V19_tmp;
                
                
                
                float V21_tmp = exp(V20_tmp);
                V21[offset2] = V21_tmp;
                
                
                
                V22_tmp += V21_tmp;
                
                
            }
            
            
            V22[offset0] = V22_tmp;
            
            
            
        }
    }
}extern "C" __global__ void K5
(float *V21, float *V22, float *Vfeat_srcinb, float *V25, 
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
            
            float V25_tmp = 0;
            int offset0 = dst_id * 8 + tx/16;int offset3 = dst_id * 128 + tx;
            
            for (int e=beg;e<end;++e) {
                
                int src_id = __ldg(column_indices + e);
                int eid = __ldg(eids + e);
                
                int offset2 = src_id * 128 + tx;int offset1 = eid * 8 + tx/16;
                
                
                
                float V23_tmp = V21[offset1]/V22[offset0];
                
                
                
                float V24_tmp = V23_tmp*Vfeat_srcinb[offset2];
                
                
                
                
                V25_tmp += V24_tmp;
                
                
            }
            
            
            V25[offset3] = V25_tmp;
            
            
            
        }
    }
}extern "C" __global__ void K6
(float *V21, float *V22, float *V25, float *V26, float *Velinb, float *Vercen, float *Vfeat_srcinb, float *V31, float *V47, float *V49, 
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
            
            float V47_tmp = 0;float V49_tmp = 0;float V31_tmp = 0;
            int offset3 = src_id * 8 + tx/16;int offset4 = src_id * 128 + tx;
            
            for (int e=beg;e<end;++e) {
                
                int dst_id = __ldg(column_indices + e);
                int eid = __ldg(eids + e);
                
                int offset0 = dst_id * 8 + tx/16;int offset2 = dst_id * 128 + tx;int offset1 = eid * 8 + tx/16;
                
                
                
                float V23_tmp = V21[offset1]/V22[offset0];
                
                
                
                float V30_tmp = V26[offset2]*V23_tmp;
                
                
                
                float V18_tmp = Velinb[offset3] + Vercen[offset0];
                
                
                
                float V19_tmp = V18_tmp - V18_tmp;
                
                
                
                float V28_tmp = V26[offset2]*Vfeat_srcinb[offset4];
                
                
                
                float V32_tmp = 1/V22[offset0];
                
                
                
                float V33_tmp = V28_tmp*V32_tmp;
                
                
                
                float V34_tmp = V26[offset2]/V22[offset0];
                
                
                
                float V35_tmp = V34_tmp*V25[offset2];
                
                
                
                float V36_tmp = -1*V35_tmp;
                
                
                
                float V40_tmp = V33_tmp + V36_tmp;
                
                
                
                float V41_tmp = V40_tmp*V21[offset1];
                
                
                
                float V42_tmp = V19_tmp>0?1:0.2;
                
                
                
                float V43_tmp = V41_tmp*V42_tmp;
                
                
                
                
                V47_tmp += V43_tmp;
                
                
                V49_tmp = V43_tmp;
                atomicAdd(V49+offset0, V49_tmp);
                
                V31_tmp += V30_tmp;
                
                
            }
            
            
            atomicAdd(V47+offset3, V47_tmp);
            
            
            
            V31[offset4] = V31_tmp;
            
            
            
        }
    }
}