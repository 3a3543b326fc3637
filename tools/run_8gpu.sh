set -x
for i in 0 1 2 3 4 5 6 7; do CUDA_VISIBLE_DEVICES=$i timeout 120 python tools/pcie_duplex.py 256 > gpurun_out/r2p_pcie_gpu$i.txt 2>&1 & done; wait
timeout 120 python tools/pcie_duplex.py 256 > gpurun_out/r2p_pcie_alone.txt 2>&1
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2p_bench_8gpu.json 2> gpurun_out/r2p_bench_8gpu.err; echo rc=$?
nvidia-smi topo -m > gpurun_out/r2p_topo.txt 2>&1; lscpu | head -25 > gpurun_out/r2p_lscpu.txt
tail -c 300 gpurun_out/r2p_bench_8gpu.json; head -3 gpurun_out/r2p_pcie_gpu0.txt; head -3 gpurun_out/r2p_pcie_alone.txt
