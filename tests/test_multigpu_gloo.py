"""CPU, world_size 2, gloo: the host-side logic of the multi-GPU path (physically-based-rendering_b200/
multigpu.py) -- partitioning, seed schedule, the one collective per frame -- with the oracle standing
in for the device renderer.  Tiles must reproduce the single-process frame bit for bit; sample sharding
must equal the mean of the per-rank running averages."""
import os
import socket
import sys

import numpy as np
import pytest

import helpers as Hh

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

W, H, FRAMES = 48, 40, 3


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _prepared():
    from oracle import oracle as O
    scene = O.load_obj(Hh.model_path("suzanne.obj"), 0)
    return Hh.Prepared(scene, W, H, max_depth=3)


def _worker(rank, world, port, mode, out_dir):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import pbr_b200
    from pbr_b200 import multigpu
    from oracle import oracle as O
    from oracle import scene as S
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = _prepared()
    img = np.zeros((H, W, 4), np.float32)
    shown = None
    for k in range(FRAMES):
        if mode == "tiles":
            y0, y1 = multigpu.tile_rows(H, rank, world)
            part, _, _ = O.path_tracing(p.defines, multigpu.frame_seed(k), multigpu.pixel_weight(k), p.px_dim, p.camera,
                                        p.nodes, p.facesV, p.facesN, p.vertices4, p.normals4, p.materials, p.lights,
                                        img, y0=y0, y1=y1, nthreads=1)
            t = torch.from_numpy(part)
            multigpu.combine_tiles(t, rank, world)          # ONE collective per frame
            img = t.numpy().copy()
            shown = img
        elif mode == "stripes":
            # what a rank holds after rendering with pbr_set_tile_stripes: its own stripes of the frame, the rest stale
            stripe = multigpu.stripe_rows_for(H, world, want=5)
            full, _, _ = O.path_tracing(p.defines, multigpu.frame_seed(k), multigpu.pixel_weight(k), p.px_dim, p.camera,
                                        p.nodes, p.facesV, p.facesN, p.vertices4, p.normals4, p.materials, p.lights,
                                        img, nthreads=1)
            mine = ((np.arange(H) // stripe) % world) == rank
            part = np.where(mine[:, None, None], full, np.float32(-7.0)).astype(np.float32)
            t = torch.from_numpy(part)
            multigpu.combine_stripes(t, rank, world, stripe)  # ONE collective per frame
            img = t.numpy().copy()
            shown = img
        else:
            g = multigpu.global_frame_index(k, rank, world)
            img, _, _ = O.path_tracing(p.defines, multigpu.frame_seed(g), multigpu.pixel_weight(k), p.px_dim, p.camera,
                                       p.nodes, p.facesV, p.facesN, p.vertices4, p.normals4, p.materials, p.lights,
                                       img, nthreads=1)
            shown = multigpu.combine_spp(torch.from_numpy(img), world).numpy()   # ONE collective per frame
    np.save(os.path.join(out_dir, "%s_rank%d.npy" % (mode, rank)), shown)
    np.save(os.path.join(out_dir, "%s_local%d.npy" % (mode, rank)), img)
    dist.destroy_process_group()


def _run(mode, tmp_path, world=2):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, mode, str(tmp_path)), nprocs=world, join=True)
    return [np.load(tmp_path / ("%s_rank%d.npy" % (mode, r))) for r in range(world)], \
        [np.load(tmp_path / ("%s_local%d.npy" % (mode, r))) for r in range(world)]


def test_tile_rows_cover_the_image():
    from pbr_b200 import multigpu
    for height in (40, 600, 1080, 2160, 37):
        for world in (1, 2, 4, 8):
            rows = [multigpu.tile_rows(height, r, world) for r in range(world)]
            assert rows[0][0] == 0 and rows[-1][1] == height
            assert all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
            assert all(y0 % 4 == 0 for y0, _ in rows)
            sizes = [y1 - y0 for y0, y1 in rows]
            assert max(sizes) - min(sizes) <= 4 or height < 4 * world


def test_seed_schedule_is_disjoint():
    from pbr_b200 import multigpu
    seen = set()
    for r in range(8):
        for k in range(16):
            g = multigpu.global_frame_index(k, r, 8)
            assert g not in seen
            seen.add(g)
    assert seen == set(range(128))
    assert multigpu.frame_seed(0) == np.float32(0.0333) and multigpu.pixel_weight(3) == np.float32(0.75)


def test_tile_sharding_is_bit_identical_to_one_process(tmp_path):
    shown, _ = _run("tiles", tmp_path)
    want, _, _ = _prepared().oracle_frames(FRAMES, nthreads=2)
    for img in shown:
        assert Hh.images_equal(img, want)


def test_stripe_sharding_is_bit_identical_to_one_process(tmp_path):
    from pbr_b200 import multigpu
    assert multigpu.stripe_rows_for(H, 2, want=5) > 0 and multigpu.stripe_rows_for(2160, 8, want=8) == 6
    shown, _ = _run("stripes", tmp_path)
    want, _, _ = _prepared().oracle_frames(FRAMES, nthreads=2)
    for img in shown:
        assert Hh.images_equal(img, want)


def test_sample_sharding_is_the_mean_of_the_ranks(tmp_path):
    shown, local = _run("spp", tmp_path)
    assert Hh.images_equal(shown[0], shown[1])
    mean = ((local[0].astype(np.float32) + local[1].astype(np.float32)) * np.float32(0.5)).astype(np.float32)
    assert Hh.images_equal(shown[0], mean)
    # the two ranks rendered different samples
    assert not Hh.images_equal(local[0], local[1])
    # and the combination is a 2x better estimate of the same picture: close to a 6-frame single render
    p = _prepared()
    ref, _, _ = p.oracle_frames(2 * FRAMES, nthreads=2)
    assert Hh.mean_relative_error(shown[0], ref) < 0.5


def test_c_abi_row_partition_is_the_python_one():
    """pbr_tile_rows (what pbr_frame_combine(ROWS) and PathTracer::setRanks use) == multigpu.tile_rows."""
    from pbr_b200 import capi, multigpu
    for height in (1, 3, 4, 7, 100, 512, 1080, 2160):
        for world in (1, 2, 3, 4, 5, 8):
            rows = [capi.Device.tileRows(height, r, world) for r in range(world)]
            assert rows == [multigpu.tile_rows(height, r, world) for r in range(world)]
            assert rows[0][0] == 0 and rows[-1][1] == height
            assert all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
    with pytest.raises(capi.PbrError):
        capi.Device.tileRows(100, 3, 3)


def test_nccl_is_loaded_on_demand_and_an_id_can_be_made_without_a_gpu():
    """pbr_comm_unique_id dlopens libnccl.so.2: 128 bytes, different every time (no GPU, no communicator needed)."""
    from pbr_b200 import capi, host
    a, b = capi.Device.commUniqueId(), host.comm_unique_id()
    assert len(a) == 128 and len(b) == 128 and a != b


def test_checkpoint_files_are_written_atomically(tmp_path):
    from pbr_b200 import host
    img = np.arange(4 * 3 * 4, dtype=np.float32).reshape(3, 4, 4)
    path = tmp_path / "acc.bin"
    host.write_checkpoint(str(path), img, 7)
    got, sc = host.read_checkpoint(str(path), 4, 3)
    assert sc == 7 and np.array_equal(got, img) and not (tmp_path / "acc.bin.tmp").exists()
    # a target that cannot be written reports an error and leaves the good file alone
    with pytest.raises(host.HostError):
        host.write_checkpoint(str(tmp_path / "no_such_dir" / "acc.bin"), img, 8)
    got, sc = host.read_checkpoint(str(path), 4, 3)
    assert sc == 7
    host.write_pfm(str(tmp_path / "img.pfm"), img)
    assert (tmp_path / "img.pfm").read_bytes().startswith(b"PF\n4 3\n-1.0\n")
