"""Stand-in for the reference's visualize_q.py (itself a no-op stub there): train_biear.py:14 imports it."""


def visualize_Q_LR(model, dataloader, device, save_dir, max_batches=5, sample_per_batch=1):
    return None
