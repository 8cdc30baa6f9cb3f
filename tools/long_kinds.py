#!/usr/bin/env python3
"""bpe_long device time for batches made of ONE kind of long piece each (a measurement aid)."""
import os; os.environ.setdefault("CFBPE_ALLOW_STAND_IN", "1")   # measurement aids run on the stand-in vocabularies
import os, sys, json, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cyberfabric-core_b200")):
    sys.path.insert(0, p)
import numpy as np, torch
from cfbpe import plugin as P, _native as N

plug = P.GpuBpeTokenizerPlugin(0, ("cl100k_base",), 64 << 20, 1 << 16)
dev = torch.device("cuda:0")
rng = random.Random(1)
L52 = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ"
def kinds(n):
    return {
        "rand52": lambda: "".join(rng.choice(L52) for _ in range(n)),
        "rand26": lambda: "".join(rng.choice(L52[:26]) for _ in range(n)),
        "uniform_a": lambda: "a" * n, "uniform_sp": lambda: " " * n, "uniform_nl": lambda: "\n" * n, "uniform_bang": lambda: "!" * n,
        "period3": lambda: ("".join(rng.choice(L52) for _ in range(3)) * n)[:n],
        "period2": lambda: ("".join(rng.choice(L52) for _ in range(2)) * n)[:n],
        "period4": lambda: ("".join(rng.choice(L52) for _ in range(4)) * n)[:n],
        "cjk": lambda: "".join(rng.choice("的一是在不了有和人这中大为上个国我以要他时来用们生到作地于出就分对成会可主发年动同工也能下过子说产种面而方后多定行学法所民得经十三之进着等部度家电力里如水化高自二理起小物现实加量都两体制机当使点从业本去把性好应开它合还因由其些然前外天政四日那社义事平形相全表间样与关各重新线内数正心反你明看原又么利比或但质气第向道命此变条只没结解问意建月公无系军很情者最立代想已通并提直题党程展五果料象员革位入常文总次品式活设及管特件长求老头基资边流路级少图山统接知较将组见计别她手角期根论运农指几九区强放决西被干做必战先回则任取据处队南给色光门即保治北造百规热领七海口东导器压志世金增争济阶油思术极交受联什认六共权收证改清己美再采转更单风切打白教速花带安场身车例真务具万每目至达走积示议声报斗完类八离华名确才科张信马节话米整空元况今集温传土许步群广石记需段研界拉林律叫且究观越织装影算低持音众书布复容儿须际商非验连断深难近矿千周委素技备半办青省列习响约支般史感劳便团往酸历市克何除消构府称太准精值号率族维划选标写存候毛亲快效斯院查江型眼王按格养易置派层片始却专状育厂京识适属圆包火住调满县局照参红细引听该铁价严") for _ in range(n // 3)),
    }
for n in (4096, 1024, 128):
    for name, gen in kinds(n).items():
        cnt = {4096: 800, 1024: 3000, 128: 50000}[n]
        texts = [gen() for _ in range(cnt)]
        data, offs = P.pack_texts(texts)
        total, np_ = int(offs[-1]), len(texts)
        d_bytes = torch.zeros(total + 256, dtype=torch.uint8, device=dev); d_bytes[:total] = torch.from_numpy(data.copy()).to(dev)
        d_offs = torch.from_numpy(offs.astype(np.int64)).to(dev)
        d_ids = torch.empty(total + 1, dtype=torch.int32, device=dev); d_off = torch.zeros(np_ + 1, dtype=torch.int64, device=dev); d_cnt = torch.empty(np_, dtype=torch.int32, device=dev)
        s = torch.cuda.current_stream().cuda_stream
        plug.ctx.profile_enable(True)
        ms = []
        for i in range(4):
            nt = plug.ctx.encode_batch_device(np_, d_bytes.data_ptr(), total, d_offs.data_ptr(), None, d_ids.data_ptr(), d_ids.numel(), d_off.data_ptr(), d_cnt.data_ptr(), s, sync=True)
            pr = plug.ctx.profile_read(); ms.append(pr["kernel_ms"]["bpe_long"]); ls = locals().setdefault("ls", []); ls.append(pr["kernel_ms"]["bpe_list"])
        print(json.dumps({"kind": name, "piece_bytes": n, "pieces": cnt, "long": pr["n_long_pieces"], "tokens": nt, "bpe_long_ms": round(min(ms[1:]), 3), "bpe_list_ms": round(min(ls[-3:]), 3), "list_pieces": pr["n_list_pieces"],
                          "split_ms": round(pr["kernel_ms"]["pretok_split"], 3), "encode_ms": round(pr["kernel_ms"]["bpe_encode"], 3)}), flush=True)
