"""previewer::infer_insertsize (meta/previewer.cc:151-304) on the C ABI: the record loop on the host (host/packer.cc:
packer_preview_add), previewer::process for all preview bundles at once on the device (agpu_batch_coverage_edit +
agpu_batch_preview), the histogram and its percentiles on the host (packer_insertsize_profile)."""
from . import hostlib as H


def infer_insertsize(ctx, sample, packer_params, gpu_params, max_preview_reads=2000000, min_preview_spliced_reads=100, min_num_hits_in_bundle=10):
    """sample: decoded coordinate-sorted records of one BAM (hostlib.Synth.sample / hostlib.read_bam layout).  Returns the
    sample_profile fields the reference's previewer fills: insert_total, insertsize_low / high / median / ave / std."""
    batch, event, skip, extra = H.preview_pack(sample, packer_params, min_num_hits_in_bundle)
    if batch.n_bundles == 0:
        return H.insertsize_profile(batch.a["bundle_hit_off"][:1] * 0, batch.a["pos"][:0], event, max_preview_reads, min_preview_spliced_reads)
    bt = ctx.upload(batch.view(), keepalive=batch)
    try:
        bt.coverage_edit(skip, extra)
        clu_off, isize = bt.preview(gpu_params)
    finally:
        bt.free()
    return H.insertsize_profile(clu_off, isize, event, max_preview_reads, min_preview_spliced_reads)
