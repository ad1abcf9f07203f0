"""BASELINE.json configs 4 and 5 (reported numbers, not bench lines).

  config 4: single-encoder SigLIP baseline, B = 256 (projector + spatial pool only, E = 1 mix is the identity)
  config 5: merv-full end-to-end generate: fusion prefix feeding a random-init Llama-2-7B prefill; time-to-visual-prefix
            and TTFT per video (synthetic input_ids, no tokenizer offline).  The LLM is a library consumer (HF transformers).
"""
import json, os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
import merv_b200 as M

dev = "cuda:0"
report = {}


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


# ---- config 4 ----
B = 256
m = M.MervFusion.build([768], 4096, [16], 64, "linear", seed=768).to(device=dev, dtype=torch.bfloat16).eval().requires_grad_(False)
g = torch.Generator(device=dev).manual_seed(1)
x = torch.randn((B, 16, 196, 768), generator=g, device=dev).to(torch.bfloat16)
with torch.inference_mode():
    ms = timeit(lambda: m([x]))
    out, w = m([x])
flops = 2 * 1024 * 4096 * 768 * B
byts = (x.numel() + out.numel()) * 2
report["config4_siglip_single_B256"] = {"ms": ms, "videos_per_s": B / ms * 1e3, "TFLOPs": flops / ms / 1e9, "compulsory_GBps": byts / ms / 1e6,
                                        "weights_all_one": bool((w.float() == 1).all())}
print("config 4:", report["config4_siglip_single_B256"], flush=True)
del x, out, m
torch.cuda.empty_cache()

# ---- config 5 ----
try:
    from transformers import LlamaConfig, LlamaForCausalLM

    cfg = LlamaConfig()  # defaults = Llama-2-7B (4096 / 32 layers / 32 heads / 11008 / 32000)
    cfg.vocab_size = 32064  # the reference pads the vocabulary to a multiple of 64 (llama2.py:74-76)
    with torch.device("meta"):
        llm = LlamaForCausalLM(cfg)
    llm = llm.to_empty(device=dev).to(torch.bfloat16)
    with torch.no_grad():
        for p_ in llm.parameters():
            p_.normal_(0, 0.02)
    llm.eval()
    fusion = M.MervFusion.build([1024, 1024, 768, 768], 4096, [16] * 4, 64, "linear", seed=1024).to(device=dev, dtype=torch.bfloat16).eval().requires_grad_(False)
    feats = [torch.randn((1, 16, n, c), generator=g, device=dev).to(torch.bfloat16) for n, c in zip([256, 256, 196, 196], [1024, 1024, 768, 768])]
    ids = torch.randint(0, 32000, (1, 33), device=dev)  # BOS + 32 text tokens

    def prefix_only():
        return fusion(feats)

    def ttft():
        emb = llm.get_input_embeddings()(ids)
        buf, _ = fusion.forward_into_embeddings(feats, emb, bos_token_length=1)  # [BOS | 1024 prefix | text], prefix written in place
        logits = llm(inputs_embeds=buf, use_cache=True).logits[:, -1]
        return logits.argmax(-1)

    with torch.inference_mode():
        t_prefix = timeit(prefix_only, n=50)
        t_ttft = timeit(ttft, n=5, warm=2)
    report["config5_generate_B1"] = {"time_to_visual_prefix_ms": t_prefix, "ttft_ms": t_ttft, "prefill_tokens": 1 + 1024 + 32,
                                     "llm": "random-init LlamaForCausalLM(LlamaConfig()) 7B bf16, HF transformers sdpa"}
    print("config 5:", report["config5_generate_B1"], flush=True)
except Exception as e:  # the LLM is only a timing sink
    report["config5_exception"] = repr(e)
    print("config 5 failed:", e, flush=True)
os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
json.dump(report, open(os.path.join(REPO, "gpurun_out", "configs45.json"), "w"), indent=1)
