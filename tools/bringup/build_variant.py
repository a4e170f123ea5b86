import importlib.util,sys,os
tag, flags = sys.argv[1], sys.argv[2]
os.environ["WEALY_NVCC_EXTRA"]=flags
spec=importlib.util.spec_from_file_location("b","/root/repo/audio-based-lyrics-matching_b200/build.py"); b=importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
b.OUT=b.OUT.replace("libwealy_b200.so","libwealy_b200_%s.so"%tag); b.STAMP=b.OUT+".srchash"
print(b.build(force=True))
