"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Imports the *unmodified* reference (nv-nguyen/nope, mounted read-only at
/root/reference) in a container that lacks some of its third-party imports.
The missing packages are replaced by minimal in-memory stubs (nothing on the
hot path uses them: they are Lightning plumbing, plotting and rendering).

Only `oracle/make_golden.py` and `bench.py --impl reference` (when the
reference tree is present) use this.  /root/reference does not exist on the
GPU box, so nothing in tests -m gpu / smoke() / the default bench arm may call
`import_reference()`.

Stubbed:
  pytorch_lightning (LightningModule -> nn.Module), diffusers, einops_exts,
  pytorch3d.transforms, imageio, trimesh, pyrender, matplotlib(.pyplot),
  moviepy, ruamel.yaml, omegaconf.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("NOPE_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "model"))


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []  # behave like a package so "import a.b" works
    sys.modules[name] = m
    parent, _, child = name.rpartition(".")
    if parent and parent in sys.modules:
        setattr(sys.modules[parent], child, m)
    return m


def _install_stubs():
    import torch
    from torch import nn

    def _missing(name):
        try:
            __import__(name)
            return False
        except Exception:
            return True

    if _missing("pytorch_lightning"):
        class LightningModule(nn.Module):
            """nn.Module with the few attributes the reference touches."""
            global_step = 0
            global_rank = 0

            @property
            def device(self):
                try:
                    return next(self.parameters()).device
                except StopIteration:
                    return torch.device("cpu")

            @property
            def dtype(self):
                try:
                    return next(self.parameters()).dtype
                except StopIteration:
                    return torch.float32

            def log(self, *a, **k):
                pass

        def seed_everything(seed, workers=False):
            torch.manual_seed(seed)

        _mod("pytorch_lightning", LightningModule=LightningModule,
             seed_everything=seed_everything)

    if _missing("diffusers"):
        class AutoencoderKL(nn.Module):
            pass
        _mod("diffusers", AutoencoderKL=AutoencoderKL)

    if _missing("einops_exts"):
        import einops

        def _many(fn):
            def inner(tensors, pattern, **kw):
                return (fn(t, pattern, **kw) for t in tensors)
            return inner

        def check_shape(t, pattern, **kw):
            return einops.rearrange(t, f"{pattern} -> {pattern}", **kw)

        _mod("einops_exts", rearrange_many=_many(einops.rearrange),
             repeat_many=_many(einops.repeat), check_shape=check_shape)

    if _missing("pytorch3d"):
        def so3_relative_angle(R1, R2, cos_angle=False, cos_bound=1e-4, eps=1e-4):
            # restated: angle of R1 R2^T from its trace
            R12 = torch.bmm(R1, R2.transpose(1, 2))
            cos = ((R12[:, 0, 0] + R12[:, 1, 1] + R12[:, 2, 2]) - 1.0) * 0.5
            if cos_angle:
                return cos
            return torch.acos(cos.clamp(-1 + cos_bound, 1 - cos_bound))

        _mod("pytorch3d")
        tr = _mod("pytorch3d.transforms", so3_relative_angle=so3_relative_angle)
        sys.path.insert(0, REFERENCE_ROOT)
        try:
            from src.poses import rotation_conversions as rc
            for n in dir(rc):
                if not n.startswith("_") and not hasattr(tr, n):
                    setattr(tr, n, getattr(rc, n))
        except Exception:
            pass

    for name in ("imageio", "trimesh", "pyrender", "moviepy", "moviepy.video",
                 "moviepy.video.io", "moviepy.video.io.bindings", "ruamel",
                 "ruamel.yaml", "omegaconf", "omegaconf.listconfig"):
        if _missing(name):
            _mod(name)
    lc = sys.modules.get("omegaconf.listconfig")
    if lc is not None and not hasattr(lc, "ListConfig"):
        lc.ListConfig = type("ListConfig", (list,), {})   # openaimodel.py:495 only type-checks it
    if "moviepy.video.io.bindings" in sys.modules and not hasattr(
            sys.modules["moviepy.video.io.bindings"], "mplfig_to_npimage"):
        sys.modules["moviepy.video.io.bindings"].mplfig_to_npimage = lambda f: None
    if _missing("matplotlib"):
        _mod("matplotlib", use=lambda *a, **k: None)
        _mod("matplotlib.pyplot")
        _mod("matplotlib.cm")
        _mod("matplotlib.patches")


_imported = None


def import_reference():
    """Returns a namespace with the reference's hot-path classes."""
    global _imported
    if _imported is not None:
        return _imported
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from src.model.u_net.denoising_diffusion_pytorch.u_net import UNet
    from src.model.encoder.template import FeatureExtractor
    from src.model.model import PoseConditional
    from src.poses import rotation_conversions
    ns = types.SimpleNamespace(
        UNet=UNet, FeatureExtractor=FeatureExtractor,
        PoseConditional=PoseConditional,
        rotation_conversions=rotation_conversions)
    _imported = ns
    return ns


def build_reference_model(u_net_dim=192, descriptor_size=8, save_dir="/tmp/nope_ref_out"):
    """Reference PoseConditional wired as configs/model/template_base.yaml:1-27."""
    ref = import_reference()
    enc = ref.FeatureExtractor(descriptor_size=descriptor_size, threshold=0.2,
                               normalize=False)
    unet = ref.UNet(u_net_dim=u_net_dim, rot_representation_dim=6, encoder=enc,
                    pose_mlp_name="single_layer")
    optim = types.SimpleNamespace(lr=5e-5, weight_decay=5e-4, warm_up_steps=500,
                                  use_inv_deltaR=True, loss_type="l1")
    testing = types.SimpleNamespace(similarity_metric="l2")
    model = ref.PoseConditional(u_net=unet, optim_config=optim,
                                testing_config=testing, save_dir=save_dir)
    return model.eval()


def build_reference_ldm_unet(model_channels=256, channel_mult=(1, 2, 4), context_dim=512,
                             attention_resolutions=(4, 2, 1), num_res_blocks=2,
                             num_head_channels=32):
    """Reference UNetModelPose wired as configs/model/vae_cin_ldm.yaml:2-31 (encoder stubbed)."""
    import torch
    import_reference()
    from src.model.u_net.ldm.adapt_openaimodel import UNetModelPose

    class _LatentStub(torch.nn.Module):
        # Stands in for VAE_StableDiffusion (src/model/encoder/AutoencoderKL.py:16-47, diffusers
        # 0.14.0 AutoencoderKL: absent here): only `latent_dim` / `name` are read by
        # UNetModelPose (adapt_openaimodel.py:123-125).  The LDM variant runs on latents.
        latent_dim = 4
        name = "VAE"

    m = UNetModelPose(
        injecting_condition_twice=False, pose_mlp_name="single_layer", rot_representation_dim=6,
        encoder=_LatentStub(), image_size=32, in_channels=4, model_channels=model_channels,
        out_channels=4, num_res_blocks=num_res_blocks,
        attention_resolutions=list(attention_resolutions), channel_mult=list(channel_mult),
        num_head_channels=num_head_channels, use_spatial_transformer=True, transformer_depth=1,
        context_dim=context_dim)
    return m.eval()
