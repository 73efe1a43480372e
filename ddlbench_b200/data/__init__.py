

def make_source(cfg, batch_size, device, rank=0, world=1, split="train", num_batches=4,
                pinned=False):
    """The training/validation batch source for a run: the on-disk image tree when real data
    was requested (``run.sh -s``, ``DATADIR`` set), otherwise in-memory synthetic batches."""
    import os
    from .synthetic import SyntheticImages
    if not cfg.synthetic or cfg.data_mode == "jpeg":
        from .imagefolder import ArrayFolder, JpegFolder, has_native_format
        if not cfg.datadir or not os.path.isdir(cfg.datadir):
            raise FileNotFoundError("real-data mode needs DATADIR (got %r)" % cfg.datadir)
        if has_native_format(cfg.datadir, cfg.benchmark):
            # the real MNIST / CIFAR-10 files as torchvision's dataset classes read them
            return ArrayFolder(cfg.datadir, cfg.benchmark, batch_size, split=split,
                               shuffle=(split == "train"), seed=cfg.seed, rank=rank, world=world,
                               workers=cfg.cores_gpu or None, pinned=True)
        # CORES_GPU is the reference's DataLoader worker count per GPU (run_template.sh:73): the same
        # number of decode worker processes here (DDLB_DECODE_PROCS overrides; 0 = thread pool only)
        procs = int(os.environ["DDLB_DECODE_PROCS"]) if "DDLB_DECODE_PROCS" in os.environ else \
            (cfg.cores_gpu if (cfg.cores_gpu or 0) >= 2 and split == "train" else 0)
        return JpegFolder(cfg.datadir, cfg.benchmark, batch_size, split=split,
                          shuffle=(split == "train"), seed=cfg.seed, rank=rank, world=world,
                          workers=cfg.cores_gpu or None, pinned=True,
                          gpu_decode=os.environ.get("DDLB_GPU_JPEG", "0") == "1", device=device,
                          decode_procs=procs)
    return SyntheticImages(cfg.benchmark, batch_size, num_batches=num_batches,
                           device=("cpu" if pinned else device), pinned=pinned,
                           seed=cfg.seed + rank)
