"""Dev aid: per-step error of the CUDA Tacotron decoder vs the fp32 and fp64 oracle on the shipped checkpoint."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np
from oracle import tacotron_oracle as to
from taco_common import real_taco_weights, sentences
from tacotronv2_wavernn_chinese_b200.tacotron.engine import TacoDecoderEngine
w = real_taco_weights(); ids = sentences()['sentences']['241']['ids']
mem = to.encoder(w, ids)
masks = (np.random.RandomState(1238).uniform(size=(1, 700, 2, 256)) >= 0.5).astype(np.uint8)
r32 = to.decode(w, mem, dropout_masks=masks[0], max_iters=200)
to.F32 = np.float64
r64 = to.decode({k: v.astype(np.float64) for k, v in w.items()}, mem.astype(np.float64), dropout_masks=masks[0], max_iters=200)
to.F32 = np.float32
out = TacoDecoderEngine(w).decode(mem[None], masks=masks, max_steps=700)
fr = out['frames'].cpu().numpy()[0]
for s in (0, 1, 2, 3, 5, 10, 20, 40, 60, 79, 100, 150):
    print(s, 'gpu-f32 %.2e  gpu-f64 %.2e  f32-f64 %.2e' % (np.abs(fr[s] - r32['frames'][s]).max(), np.abs(fr[s] - r64['frames'][s]).max(), np.abs(r32['frames'][s] - r64['frames'][s]).max()),
          'stop gpu %.5f f64 %.5f' % (out['stop'][0, s].item(), r64['stop'][s]), 'align err %.2e' % np.abs(out['align'][0, s].cpu().numpy() - r64['alignments'][s]).max())
