import cProfile, pstats, sys, io, os
sys.argv = ["bench_train.py", "--steps", "10", "--warmup", "3"]
sys.path.insert(0, os.path.join(os.getcwd(), "tools"))
import importlib.util
spec = importlib.util.spec_from_file_location("bt", "tools/bench_train.py")
bt = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bt)
pr = cProfile.Profile()
pr.enable()
bt.main()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(s.getvalue()[:9000])
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(25)
print(s.getvalue()[:5000])
