"""cProfile of the θ-sharded driver on rank 0 (host-side overhead of the replicated control flow):
torchrun --nproc-per-node G tools/prof_dist.py c5"""
import cProfile, io, os, pstats, runpy, sys
rank = int(os.environ.get("RANK", 0))
pr = cProfile.Profile()
sys.argv = ["tools/smc2_dist.py"] + sys.argv[1:]
pr.enable()
try:
    runpy.run_path("tools/smc2_dist.py", run_name="__main__")
finally:
    pr.disable()
    if rank == 0:
        s = io.StringIO()
        st = pstats.Stats(pr, stream=s)
        st.sort_stats("cumtime").print_stats(r"sequential_monte_carlo_b200|smc2_dist", 40)
        st.sort_stats("tottime").print_stats(25)
        open("gpurun_out/prof_dist_rank0.txt", "w").write(s.getvalue())
