"""Frame sharding across the GPUs of one box and the single collective of the path.

The hot path is stateless per frame (CAPE) / per (initial pose, match list) (pose solve), so a batch of frames is split
into contiguous shards, one per rank, with NO data-path collective; the only exchange is one all-gather of the per-frame
poses [frames x 7] FP64 (position + unit quaternion) so that every rank holds all poses. Backend: NCCL over NVLink on
GPUs, gloo in the CPU tests. Shards may be uneven (n_frames not divisible by world): the gather pads to the largest."""
import torch
import torch.distributed as dist


def frame_shard(n_frames, rank, world):
    """Contiguous [start, stop) of `rank`; the first n_frames % world ranks get one extra frame."""
    if world <= 0 or not (0 <= rank < world) or n_frames < 0:
        raise ValueError("bad shard request: n_frames=%r rank=%r world=%r" % (n_frames, rank, world))
    base, extra = divmod(n_frames, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_sizes(n_frames, world):
    return [frame_shard(n_frames, r, world)[1] - frame_shard(n_frames, r, world)[0] for r in range(world)]


def gather_poses(local_poses, n_frames, group=None):
    """All-gather of the per-frame poses. local_poses: [local_frames, 7] float64 tensor of this rank's shard (on the
    device the process group's backend expects). Returns [n_frames, 7] in global frame order on every rank."""
    if not (dist.is_available() and dist.is_initialized()):
        if local_poses.shape[0] != n_frames:
            raise ValueError("no process group: the local shard must be the whole batch")
        return local_poses.clone()
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = shard_sizes(n_frames, world)
    if local_poses.shape[0] != sizes[rank] or local_poses.shape[1:] != (7,):
        raise ValueError("rank %d holds %r, expected [%d, 7]" % (rank, tuple(local_poses.shape), sizes[rank]))
    cap = max(sizes)
    send = local_poses
    if sizes[rank] != cap:
        send = torch.zeros((cap, 7), dtype=local_poses.dtype, device=local_poses.device)
        send[:sizes[rank]] = local_poses
    recv = torch.empty((world, cap, 7), dtype=local_poses.dtype, device=local_poses.device)
    dist.all_gather_into_tensor(recv.view(-1), send.contiguous().view(-1), group=group)
    if all(s == cap for s in sizes):
        return recv.view(world * cap, 7)
    return torch.cat([recv[r, :sizes[r]] for r in range(world)], dim=0)
