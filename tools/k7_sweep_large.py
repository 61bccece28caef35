import os, subprocess, sys
ROOT = "/root/repo"
code = r'''
import os, sys
sys.path.insert(0, os.path.join(%r, "stereo-vision_b200"))
import elas_b200, synth
W,H,D = %d,%d,%d
L,R,_ = synth.synthetic_pair(W,H,D,0)
e = elas_b200.ElasB200(elas_b200.stereomapper(D), W, H, n_slots=1)
e.process(L,R)
print("%%s thr=%%s seg=%%s  %%.2f us (flushed) %%.2f us (warm)" %% ((W,H), os.environ.get("ELAS_B200_K7_THREADS"), os.environ.get("ELAS_B200_K7_SEG"), e.time_matching(30, True)*1e3, e.time_matching(30, False)*1e3))
e.close()
'''
for (W, H, D) in [(4096, 2160, 256), (1920, 1080, 128)]:
    for thr in (128, 256):
        for seg in (224, 320, 448, 640):
            env = dict(os.environ, ELAS_B200_K7_THREADS=str(thr), ELAS_B200_K7_SEG=str(seg))
            subprocess.run([sys.executable, "-c", code % (ROOT, W, H, D)], env=env)
