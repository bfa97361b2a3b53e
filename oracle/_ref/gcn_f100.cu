extern "C" __global__ void K12
(float *Vhinb, float *Vnormcen, float *Vnorminb, float *V96, 
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
            
            float V95_tmp = 0;
            int offset2 = dst_id * 1 + tx/100;int offset3 = dst_id * 100 + tx;
            
            for (int e=beg;e<end;++e) {
                
                int src_id = __ldg(column_indices + e);
                int eid = __ldg(eids + e);
                
                int offset0 = src_id * 1 + tx/100;int offset1 = src_id * 100 + tx;
                
                
                
                float V94_tmp = Vhinb[offset1]*Vnorminb[offset0];
                
                
                
                
                V95_tmp += V94_tmp;
                
                
            }
            
            
            
            
            
            
            
            float V96_tmp = V95_tmp*Vnormcen[offset2];
            V96[offset3] = V96_tmp;
            
        }
    }
}extern "C" __global__ void K13
(float *V97, float *Vnormcen, float *Vnorminb, float *V101, 
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
            
            float V100_tmp = 0;
            int offset2 = src_id * 1 + tx/100;int offset3 = src_id * 100 + tx;
            
            for (int e=beg;e<end;++e) {
                
                int dst_id = __ldg(column_indices + e);
                int eid = __ldg(eids + e);
                
                int offset0 = dst_id * 1 + tx/100;int offset1 = dst_id * 100 + tx;
                
                
                
                float V98_tmp = V97[offset1]*Vnormcen[offset0];
                
                
                
                
                V100_tmp += V98_tmp;
                
                
            }
            
            
            
            
            
            
            
            float V101_tmp = V100_tmp*Vnorminb[offset2];
            V101[offset3] = V101_tmp;
            
        }
    }
}