// removed from CCCL; pcsr.cu includes it but uses nothing of it
