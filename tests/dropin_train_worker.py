"""Worker of tests/test_gpu_dropin_reference.py: runs the reference's OWN, UNMODIFIED `Train.py` (staged under
baseline/_ref by __graft_entry__.build()) on top of this framework, the way INTEGRATION.md section 1 describes:

  * two shim files `Modules.py` / `distributed.py` (written into a scratch directory that precedes the reference
    on sys.path) re-export this package's drop-in modules;
  * the packages the reference imports for audio / plotting and that this image does not have (librosa, pysptk,
    matplotlib) are stubbed in sys.modules -- nothing on the tested path calls into them;
  * a synthetic pattern set in the reference's on-disk format (`Pattern_Generator.py:108-131`: one pickle per
    utterance with an fp16 'Mel' [80, L], METADATA.PICKLE with 'File_List_by_Speaker_Dict').

Then: Trainer.__init__ (Dataset_Generate, Model_Generate, Load_Checkpoint, Logger), Train_Step on real DataLoader
batches (Train.py:140-168), Evaluation_Epoch (Train.py:214-234, TensorBoard histograms through named_parameters),
Inference_Step (Train.py:237-242), Save_Checkpoint / Load_Checkpoint (Train.py:269-310) into a second Trainer, and the
checkpoint loaded with strict=True into the REFERENCE's own GE2E on CPU, whose d-vectors must agree.
"""
import importlib.util
import os
import pickle
import sys
import types

import numpy as np
import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def stub_modules():
    def _missing(*a, **k):
        raise RuntimeError("stubbed third-party call reached on the tested path")

    librosa = types.ModuleType("librosa")
    librosa.load = _missing
    librosa.util = types.ModuleType("librosa.util")
    librosa.util.normalize = _missing
    librosa.filters = types.ModuleType("librosa.filters")
    librosa.filters.mel = _missing
    librosa.effects = types.ModuleType("librosa.effects")
    pysptk = types.ModuleType("pysptk")
    pysptk.sptk = types.ModuleType("pysptk.sptk")
    pysptk.sptk.rapt = _missing
    mpl = types.ModuleType("matplotlib")
    mpl.use = lambda *a, **k: None
    mpl.pyplot = types.ModuleType("matplotlib.pyplot")
    mpl.colors = types.ModuleType("matplotlib.colors")
    for m in (librosa, librosa.util, librosa.filters, librosa.effects, pysptk, pysptk.sptk, mpl, mpl.pyplot, mpl.colors):
        sys.modules.setdefault(m.__name__, m)


def write_patterns(path, speakers, utts, rng):
    os.makedirs(path, exist_ok=True)
    by_speaker = {}
    for s in range(speakers):
        centre = rng.standard_normal((80, 1)) * 1.5
        files = []
        for u in range(utts):
            length = int(rng.integers(24, 70))          # some shorter than Frame_Length.Min: reflect-padded by Correction
            mel = np.clip(-5.0 + centre + rng.standard_normal((80, length)), np.log(1e-5), 2.0).astype(np.float16)
            name = "S%02d/U%02d.PICKLE" % (s, u)
            os.makedirs(os.path.join(path, "S%02d" % s), exist_ok=True)
            pickle.dump({"Mel": mel, "Speaker": "S%02d" % s}, open(os.path.join(path, name), "wb"))
            files.append(name)
        by_speaker["S%02d" % s] = files
    pickle.dump({"File_List_by_Speaker_Dict": by_speaker}, open(os.path.join(path, "METADATA.PICKLE"), "wb"))


def main():
    ref, work = sys.argv[1], sys.argv[2]
    # SPK_DROPIN_REFERENCE_MODULES=1: no shims -- the same script against the reference's own modules on CPU, which
    # is how the CPU suite checks that this harness (patterns, yaml, call sequence) is what the reference expects
    use_reference = os.environ.get("SPK_DROPIN_REFERENCE_MODULES") == "1"
    shims = os.path.join(work, "shims")
    os.makedirs(shims, exist_ok=True)
    if not use_reference:
        open(os.path.join(shims, "Modules.py"), "w").write(
            "from speaker_embedding_torch_b200.Modules import *          # noqa: F401,F403\n")
        open(os.path.join(shims, "distributed.py"), "w").write(
            "from speaker_embedding_torch_b200.distributed import (      # noqa: F401\n"
            "    init_distributed, apply_gradient_allreduce, reduce_tensor)\n")
    if os.environ.get("SPK_DROPIN_DEVICE_COLLATER") == "1" and not use_reference:
        # third shim: the training collater of this package (crop / reflect-pad on the device), everything else of
        # Datasets.py -- Dataset, Inference_Collater, Correction -- stays the reference's
        open(os.path.join(shims, "Datasets.py"), "w").write(
            "import importlib.util, os\n"
            "_s = importlib.util.spec_from_file_location('reference_Datasets', os.path.join(%r, 'Datasets.py'))\n"
            "_m = importlib.util.module_from_spec(_s); _s.loader.exec_module(_m)\n"
            "Dataset, Inference_Collater, Correction = _m.Dataset, _m.Inference_Collater, _m.Correction\n"
            "from speaker_embedding_torch_b200.Datasets import Collater      # noqa: F401\n" % ref)
    sys.path[:0] = [shims, ref, ROOT]
    stub_modules()

    rng = np.random.default_rng(0)
    write_patterns(os.path.join(work, "Train"), 8, 5, rng)
    write_patterns(os.path.join(work, "Eval"), 6, 4, rng)
    hp = yaml.load(open(os.path.join(ref, "Hyper_Parameters.yaml")), Loader=yaml.Loader)
    hp["Train"]["Train_Pattern"]["Path"] = os.path.join(work, "Train")
    hp["Train"]["Eval_Pattern"]["Path"] = os.path.join(work, "Eval")
    hp["Train"]["Batch"] = {"Train": {"Speaker": 4, "Pattern_per_Speaker": 3},
                            "Eval": {"Speaker": 3, "Pattern_per_Speaker": 3}}
    hp["Train"]["Frame_Length"] = {"Min": 40, "Max": 48}
    hp["Train"]["Inference"] = {"Samples": 5, "Frame_Length": 16, "Overlap_Length": 8}
    hp["Train"]["Learning_Rate"]["Initial"] = 1.0e-3
    hp["Checkpoint_Path"] = os.path.join(work, "Checkpoint")
    hp["Log_Path"] = os.path.join(work, "Log")
    hp_path = os.path.join(work, "Hyper_Parameters.yaml")
    yaml.dump(hp, open(hp_path, "w"))

    import Train                                         # baseline/_ref/Train.py, unmodified
    import Modules
    assert os.path.dirname(os.path.abspath(Train.__file__)) == os.path.abspath(ref)
    if not use_reference:
        import speaker_embedding_torch_b200 as pkg
        assert Modules.GE2E is pkg.GE2E and Train.GE2E is pkg.GE2E and Train.GE2E_Loss is pkg.GE2E_Loss

    torch.manual_seed(0)
    np.random.seed(0)
    import random
    random.seed(0)
    trainer = Train.Trainer(hp_path=hp_path, steps=0)
    if not use_reference:
        assert trainer.device.type == "cuda" and next(trainer.model.parameters()).is_cuda

    class _Bar:                                           # Trainer.Train() would create a tqdm here (Train.py:328-332)
        def update(self, n):
            pass
    trainer.tqdm = _Bar()
    losses = []
    it = iter(trainer.dataloader_dict["Train"])
    for _ in range(2):
        features = next(it)
        assert features.dim() == 3 and features.size(0) == 12 and features.size(1) == 80 and 40 <= features.size(2) <= 48
        trainer.Train_Step(features)
        losses.append(trainer.scalar_dict["Train"]["Loss/Embedding"])
    # a fixed batch, repeated: the loss must go down
    fixed = next(iter(trainer.dataloader_dict["Train"]))
    curve = []
    for _ in range(10):
        before = trainer.scalar_dict["Train"]["Loss/Embedding"]
        trainer.Train_Step(fixed)
        curve.append(trainer.scalar_dict["Train"]["Loss/Embedding"] - before)
    assert trainer.steps == 12 and np.isfinite(curve).all() and np.mean(curve[-3:]) < np.mean(curve[:3]), curve

    trainer.Evaluation_Epoch()                            # eval loss + parameter histograms (Logger.add_histogram_model)
    assert trainer.model.training
    feats, speakers = next(iter(trainer.dataloader_dict["Inference"]))
    trainer.model.eval()
    emb = trainer.Inference_Step(feats)                   # model(features, samples=5)
    trainer.model.train()
    assert emb.shape == (len(speakers), 256) and torch.allclose(emb.norm(dim=1), torch.ones(len(speakers), device=emb.device), atol=1e-5)

    trainer.Save_Checkpoint()
    ckpt = os.path.join(hp["Checkpoint_Path"], "S_12.pt")
    assert os.path.exists(ckpt)
    state = torch.load(ckpt, map_location="cpu")
    assert sorted(state) == ["Model", "Optimizer", "Scheduler", "Steps"] and len(state["Model"]) == 44

    trainer2 = Train.Trainer(hp_path=hp_path, steps=0)    # Load_Checkpoint picks the newest *.pt (Train.py:270-290)
    assert trainer2.steps == 12
    for (k, a), (_, b) in zip(trainer.model.state_dict().items(), trainer2.model.state_dict().items()):
        assert torch.equal(a, b), k
    trainer2.tqdm = _Bar()
    trainer2.Train_Step(fixed)                            # resumed optimiser state steps fine
    assert trainer2.steps == 13

    # the checkpoint in the REFERENCE's own module (strict), on CPU: same d-vectors
    spec = importlib.util.spec_from_file_location("reference_Modules", os.path.join(ref, "Modules.py"))
    ref_mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_mod)
    import warnings
    warnings.filterwarnings("ignore")
    ref_model = ref_mod.GE2E(Train.Recursive_Parse(hp)).eval()
    ref_model.load_state_dict(state["Model"], strict=True)
    with torch.no_grad():
        want = ref_model(feats.float(), 5)
    trainer.model.eval()
    got = trainer.Inference_Step(feats).cpu()
    cos = torch.nn.functional.cosine_similarity(got.double(), want.double(), dim=1)
    assert cos.min().item() >= 0.9999, cos.min().item()
    # and back: the reference's state_dict loads into the drop-in module
    trainer.model.load_state_dict(ref_model.state_dict(), strict=True)
    print("DROPIN_TRAIN_OK steps=%d first_losses=%s curve=%.4f->%.4f min_cos=%.7f"
          % (trainer2.steps, ["%.4f" % v for v in losses], curve[0], curve[-1], cos.min().item()))


if __name__ == "__main__":
    main()
