"""Worker of tests/test_gpu_montecarlo.py::test_two_rank_nccl_sweep_equals_the_sum_of_single_gpu_runs.

Launched with torch.distributed.run, one process per GPU: runs MonteCarloEngine.run_point on the NCCL
process group (frames of every interval sharded over the ranks, ONE all-reduce of the counters per
interval) and lets rank 0 write the reduced counters as JSON."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "ldpc-simulator_b200"))
sys.path.insert(0, HERE)


def main():
    import torch
    import torch.distributed as dist
    from conftest import load_code
    from encoder_decoder_data import EncoderDecoderData
    from mc_driver import MonteCarloEngine
    out, name, frames, interval, snr, early = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), float(sys.argv[5]), int(sys.argv[6])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    edd = EncoderDecoderData(h=load_code(name).sparse_matrix())
    eng = MonteCarloEngine(edd, graph="alist", precision="f32_fast", max_iterations=20, early_termination=bool(early),
                           fix_odd_check_sign=True, sigma_sq_quirk=False, seed=2027)
    cnt = eng.run_point(snr, 0.5, frames=frames, interval_frames=interval, random_codewords=False)
    if dist.get_rank() == 0:
        with open(out, "w") as f:
            json.dump({"world": dist.get_world_size(), "frames": cnt.frames, "frame_errors": cnt.frame_errors,
                       "bit_errors": cnt.bit_errors, "conv_sum": cnt.conv_sum, "conv_count": cnt.conv_count}, f)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
