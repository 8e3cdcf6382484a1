"""Golden vectors for the C++ byte-level BPE tokenizer (opsagent_b200/csrc/bpe.hpp): trains two small tokenizers with the Hugging Face
`tokenizers` library in the exact configuration of the Llama-3 and Qwen2.5 tokenizer.json files (Split(regex) + ByteLevel pre-tokenizer,
byte-level BPE model, special tokens) on a seeded synthetic corpus, and records what that library encodes a list of probe strings to.
    python tests/golden/gen_golden_bpe.py       -> tests/golden/bpe_{llama3,qwen2}_tiny.json, tests/golden/bpe_cases.json"""
import json
import os
import random

from tokenizers import Regex, Tokenizer, decoders, models, normalizers, pre_tokenizers, trainers

HERE = os.path.dirname(os.path.abspath(__file__))
LLAMA3 = r"(?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]+[\r\n]*|\s*[\r\n]+|\s+(?!\S)|\s+"
QWEN2 = r"(?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}| ?[^\s\p{L}\p{N}]+[\r\n]*|\s*[\r\n]+|\s+(?!\S)|\s+"


def corpus(seed=7, n=600):
    r = random.Random(seed)
    words = ["kubectl", "get", "pods", "namespace", "default", "kube-system", "Running", "CrashLoopBackOff", "image", "nginx", "error",
             "the", "cluster", "has", "namespaces", "don't", "it's", "we'll", "I'M", "THEY'RE", "集群", "命名空间", "有", "个", "节点", "错误",
             "déploiement", "über", "naïve", "Ελληνικά", "привет", "🙂", "→", "apiVersion:", "kind:", "Pod", "metadata:", "spec:", "containers:"]
    out = []
    for _ in range(n):
        k = r.randrange(3, 14)
        toks = []
        for _ in range(k):
            c = r.random()
            if c < 0.70:
                toks.append(r.choice(words))
            elif c < 0.82:
                toks.append(str(r.randrange(10 ** r.randrange(1, 7))))
            elif c < 0.90:
                toks.append(r.choice(["{", "}", "[]", "\"name\":", "--all-namespaces", "!!!", "...", "(x)", "a=b;", "#tag", "100%", "$5.00"]))
            else:
                toks.append(r.choice(["\n", "\n\n", "\t", "  ", "   \n", "\r\n"]))
        out.append(" ".join(toks))
    return out


def build(pattern, ignore_merges, specials, vocab_size, nfc=False):
    tok = Tokenizer(models.BPE(ignore_merges=ignore_merges))
    if nfc:
        tok.normalizer = normalizers.NFC()          # as in Qwen2.5's tokenizer.json
    tok.pre_tokenizer = pre_tokenizers.Sequence([pre_tokenizers.Split(Regex(pattern), behavior="isolated", invert=False),
                                                 pre_tokenizers.ByteLevel(add_prefix_space=False, use_regex=False)])
    tok.decoder = decoders.ByteLevel()
    trainer = trainers.BpeTrainer(vocab_size=vocab_size, special_tokens=specials, initial_alphabet=pre_tokenizers.ByteLevel.alphabet(), show_progress=False)
    tok.train_from_iterator(corpus(), trainer)
    return tok


PROBES = [
    "", " ", "a", "hello world", "Hello, World!", "how many namespace in the cluster?", "kubectl get pods --all-namespaces -o wide",
    "it's we'll DON'T I'M they're THEY'VE 'tis", "x=1234567890 y=12 z=007 3.14159 1,000,000", "  leading and trailing   ", "tabs\tand\nnewlines\n\nand\r\nCRLF \n ",
    "line one\n   indented\n\n\n   more", "集群中有多少个命名空间？请用kubectl查询。", "Ελληνικά привет déploiement über naïve", "emoji 🙂🙂 → arrows ←→ and ½ ² ① numbers",
    "{\"question\":\"q\",\"thought\":\"t\",\"action\":{\"name\":\"kubectl\",\"input\":\"get ns\"},\"observation\":\"\",\"final_answer\":\"\"}",
    "apiVersion: v1\nkind: Pod\nmetadata:\n  name: app-0001\n  labels:\n    tier: web\nspec:\n  containers:\n  - name: c1\n    image: nginx:1.25.3\n",
    "!!!???...;;;", "a  b   c    d", "end with space ", "\n", " \n", "\t\t", "ＡＢＣ１２３ fullwidth", "مرحبا ١٢٣", "x" * 50, "ab" * 40 + " " + "1" * 10,
]

if __name__ == "__main__":
    l3_specials = ["<|begin_of_text|>", "<|end_of_text|>", "<|start_header_id|>", "<|end_header_id|>", "<|eot_id|>"]
    q2_specials = ["<|endoftext|>", "<|im_start|>", "<|im_end|>"]
    cases = {}
    for name, pattern, ign, sp in (("llama3", LLAMA3, True, l3_specials), ("qwen2", QWEN2, False, q2_specials)):
        tok = build(pattern, ign, sp, 700, nfc=(name == "qwen2"))
        path = os.path.join(HERE, f"bpe_{name}_tiny.json")
        tok.save(path)
        cases[name] = [{"text": t, "ids": tok.encode(t, add_special_tokens=False).ids} for t in PROBES]
        for c in cases[name]:
            assert tok.decode(c["ids"], skip_special_tokens=False) == c["text"], c["text"]      # byte-level BPE round-trips
        cases[name + "_specials"] = {s: tok.token_to_id(s) for s in sp}
    json.dump(cases, open(os.path.join(HERE, "bpe_cases.json"), "w"), ensure_ascii=False, indent=0)
    print("wrote", os.listdir(HERE))
