"""Harness that drives the REFERENCE'S OWN `nerf/trainer.py::Trainer` (byte-compiled under oracle/_ref/bytecode, unmodified) on top
of a model -- this repo's drop-in `NeRFNetwork` in the GPU tests, the reference's own model in the CPU dry run that checks the
harness itself.  The trainer is the B-model caller of SURVEY.md 8b: `train_step` (:336), `eval_step` (:570), `test_step`
(:692), `post_train_step` (:558), plus main.py's optimizer wiring (`model.get_params`, main.py:283) and name-based freezing
(main.py:249-256).  Synthetic data dicts stand in for nerf/provider.py's loader (collate: provider.py:894-1114)."""
import types

import torch


def trainer_opt(**kw):
    """The `opt` fields Trainer.__init__ / the three steps read, with main.py's defaults, on top of the render fields."""
    from oracle import render_oracle as O
    opt = O.default_opt()
    extra = dict(trajectory_root=None, use_point=False, point_file=None, cache_size=0, cache_interval=4, feature_container="distill",
                 lambda_entropy=0.0, lambda_tv=1e-7, lambda_wd=1e-6, adaptive_num_rays=True, num_points=2 ** 18, num_rays=256, fp16=False,
                 val_save_root=None, render_mask_type="heatmap", render_mask_instance_id=0, epsilon=1e-6, error_map=False,
                 label_regularization_weight=0.0, ray_pair_rgb_loss_weight=0.0, ray_pair_rgb_iter=100, mixed_sampling=False,
                 num_local_sample=4, local_sample_patch_size=8, ray_pair_rgb_exp_weight=1.0, ray_pair_rgb_use_pred_logistics=False,
                 return_extra=False, workspace=None, sam_type="sam_hq", error_map_size=128, ray_pair_rgb_threshold=0.5,
                 ray_pair_rgb_num_sample=8, init_ckpt="", patch_size=1, ckpt="scratch")
    extra.update(kw)
    for k, v in extra.items():
        setattr(opt, k, v)
    return opt


def make_trainer(R, backend, model, opt, workspace, lr=1e-2, ema_decay=0.95):
    """Trainer(name, opt, model, criterion, optimizer, ema, scheduler, ...) wired the way main.py:283-304 does it."""
    import importlib
    with R.env(backend):
        T = importlib.import_module("nerf.trainer")
        if not hasattr(T.plt, "cm"):     # matplotlib stub: Trainer.__init__ builds a colour map (trainer.py:128-131)
            T.plt.cm = types.SimpleNamespace(get_cmap=lambda name, n: (lambda i: ((i % 7) / 7.0, (i % 5) / 5.0, (i % 3) / 3.0, 1.0)))
        optimizer = torch.optim.Adam(model.get_params(lr), eps=1e-15)
        scheduler = torch.optim.lr_scheduler.LambdaLR(optimizer, lambda it: 0.1 ** min(it / 100, 1))
        trainer = T.Trainer("ngp", opt, model, criterion=torch.nn.MSELoss(reduction="none"), optimizer=optimizer, ema_decay=ema_decay,
                            lr_scheduler=scheduler, device=next(model.parameters()).device, mute=True, fp16=False, workspace=workspace,
                            use_checkpoint="scratch", scheduler_update_every_step=True)
        trainer.error_map = None
    return trainer


def frame_data(H, W, pose_k, device, with_images=True, rows=None):
    """What provider.py's collate hands to eval_step / test_step for one full image."""
    from sanerf_hq_b200.rays import get_rays, lego_intrinsics, orbit_pose
    ro, rd = get_rays(orbit_pose(pose_k), lego_intrinsics(H, W), H, W)
    data = {"rays_o": ro.to(device), "rays_d": rd.to(device), "index": [pose_k], "H": H, "W": W, "img_names": None}
    if with_images:
        g = torch.Generator().manual_seed(100 + pose_k)
        data["images"] = torch.rand(H, W, 3, generator=g).to(device)
    return data


def train_data(n_rays, device, seed=0, masks=False):
    """A training batch: random pixels of random poses (rays [N,3], images [N,3]; masks [1,N] for the object stage)."""
    from sanerf_hq_b200.rays import get_rays, lego_intrinsics, orbit_pose
    g = torch.Generator().manual_seed(seed)
    ro, rd = get_rays(orbit_pose(seed % 24), lego_intrinsics(800, 800), 800, 800)
    sel = torch.randint(0, 800 * 800, (n_rays,), generator=g)
    data = {"rays_o": ro[sel].contiguous().to(device), "rays_d": rd[sel].contiguous().to(device), "index": [seed % 24], "H": 800, "W": 800,
            "images": torch.rand(n_rays, 3, generator=g).to(device)}
    if masks:
        data["masks"] = torch.randint(0, 2, (1, n_rays), generator=g).to(device)
        data["inds_coarse"] = sel.to(device)[None]
    return data


def one_optimizer_step(trainer, data):
    """The body of Trainer.train_one_epoch's loop (trainer.py:1480-1500) for one batch."""
    trainer.model.train()
    trainer.global_step += 1
    trainer.optimizer.zero_grad()
    preds, truths, loss = trainer.train_step(data)
    trainer.scaler.scale(loss).backward()
    trainer.post_train_step()
    trainer.scaler.step(trainer.optimizer)
    trainer.scaler.update()
    trainer.lr_scheduler.step()
    return preds, truths, loss
