"""Discrete-event model of the tcgen05 kernel's per-SM pipeline (loader, fold warps, MMA warp, epilogue warps), used
to rank restructurings before writing them.  Durations in SM cycles from tools/tc_trace.py timelines."""
import sys

def simulate(n=60, tL=1900, tFE=4300, tFO=4300, tM=1100, tX=350, tC=550, tF=2700, wake=120,
             fold="split", finish="after_x0", audio_bufs=1, verbose=False):
    INF = 0
    audio_full = {}; audio_empty = {-1: 0, -2: 0}
    a_fullE = {}; a_fullO = {}; a_emptyE = {-1: 0}; a_emptyO = {-1: 0}
    fe_end = {-1: 0}; fo_end = {-1: 0}
    mma_free = 0; epi_free = 0
    d_empty = 0                      # when the accumulator is free for the next unit
    tile_done = {}
    pendingF = False
    for i in range(n):
        # loader
        l_start = audio_empty[i - audio_bufs]
        audio_full[i] = l_start + tL
        # fold
        if fold == "split":          # E warps and O warps run in parallel, each the whole sweep
            s = max(audio_full[i], a_emptyE[i - 1], fe_end[i - 1]); fe_end[i] = s + tFE; a_fullE[i] = fe_end[i]
            s = max(audio_full[i], a_emptyO[i - 1], fo_end[i - 1]); fo_end[i] = s + tFO; a_fullO[i] = fo_end[i]
        else:                        # all eight warps do the E sweep, then the O sweep (half the time each)
            s = max(audio_full[i], a_emptyE[i - 1], fo_end[i - 1]); fe_end[i] = s + tFE / 2; a_fullE[i] = fe_end[i]
            s = max(fe_end[i], a_emptyO[i - 1]); fo_end[i] = s + tFO / 2; a_fullO[i] = fo_end[i]
        audio_empty[i] = max(fe_end[i], fo_end[i])
        # MMA + epilogue, unit by unit
        for u in range(4):
            ready = a_fullE[i] if u < 2 else a_fullO[i]
            m_start = max(mma_free, ready + wake, d_empty)
            d_full = m_start + tM
            mma_free = m_start + tM * 0.8          # issue is done a little before the tensor pipe
            if u == 1: a_emptyE[i] = d_full
            if u == 3: a_emptyO[i] = d_full
            x_start = max(epi_free, d_full + wake)
            x_end = x_start + tX
            d_empty = x_end + wake
            epi_free = x_end
            if finish == "after_x0" and u == 0 and pendingF:
                epi_free += tF; pendingF = False
            epi_free += tC
            if u == 3:
                if finish == "after_c3": epi_free += tF
                else: pendingF = True
                tile_done[i] = epi_free
    period = (tile_done[n - 1] - tile_done[n - 21]) / 20
    return period


# ---- round 2: the kernel as it ships (fold warps issue the tile copy themselves after the O sweep, single accumulator,
# finish after the tile's last unit) against a 3-slot operand ring with two accumulators.  Durations from
# profiles/r02_tc80_cta0_timeline.txt.  `cur` reproduces the measured ~10 k cycles per tile; the ring gains ~3 %.
def simulate_r2(mode, n=80, tL=1700, tE=3640, tO=2840, tM=1100, tX=500, tC=700, tF=3300, h=100, finish_split=False, tEscale=1.0):
    tE*=tEscale; tO*=tEscale
    mma_done={}; pull_done={}; a_full={}
    audio_full={0:tL}
    mma_free=0; epi_free=0
    E_end={}; O_end={}
    tile_done={}
    # event-driven by iterating tiles in order; dependencies only go backward except O(i) on mma_done(4i) in ring mode.
    for i in range(n):
        # E sweep
        if mode=='cur':
            dep = mma_done.get(4*(i-1)+1,0)
        else:
            dep = mma_done.get(4*(i-1)+2,0)
        sE=max(audio_full[i], dep+h, O_end.get(i-1,0))
        E_end[i]=sE+tE
        a_full[4*i]=a_full[4*i+1]=E_end[i]+h
        # MMA u0,u1 (need to interleave with epilogue) -> process units sequentially with epilogue
        def do_unit(g):
            nonlocal mma_free, epi_free
            if mode=='cur': dfree = pull_done.get(g-1,0)
            else: dfree = pull_done.get(g-2,0)
            st=max(mma_free, a_full[g], dfree+h)
            mma_done[g]=st+tM
            mma_free=st+tM*0.9
            xs=max(epi_free, mma_done[g]+h)
            pull_done[g]=xs+tX
            epi_free=pull_done[g]+tC
            if g%4==3:
                epi_free+=tF
                tile_done[g//4]=epi_free
        do_unit(4*i)
        if mode=='cur':
            depO = mma_done.get(4*(i-1)+3,0)
        else:
            depO = max(mma_done.get(4*(i-1)+3,0), mma_done[4*i])
        # u1 doesn't depend on O
        do_unit(4*i+1)
        sO=max(E_end[i], depO+h)
        O_end[i]=sO+tO
        a_full[4*i+2]=a_full[4*i+3]=O_end[i]+h
        audio_full[i+1]=O_end[i]+h+tL
        do_unit(4*i+2); do_unit(4*i+3)
    return (tile_done[n-1]-tile_done[n-21])/20


if __name__ == "__main__":
    base = dict()
    for name, kw in [
        ("current: split sweeps, finish after next X0", dict()),
        ("split sweeps, finish after C3", dict(finish="after_c3")),
        ("8-warp sweeps, finish after X0", dict(fold="all8")),
        ("8-warp sweeps, finish after C3", dict(fold="all8", finish="after_c3")),
        ("split sweeps + 2 audio buffers", dict(audio_bufs=2)),
        ("8-warp sweeps + 2 audio buffers", dict(fold="all8", audio_bufs=2)),
        ("8-warp sweeps + 2 audio buffers, finish after C3", dict(fold="all8", audio_bufs=2, finish="after_c3")),
        ("current, finish 1500", dict(tF=1500)),
        ("8-warp sweeps, finish 1500", dict(fold="all8", tF=1500)),
        ("current, sweeps 3300", dict(tFE=3300, tFO=3300)),
        ("current, MMA unit 800", dict(tM=800)),
        ("8-warp sweeps, finish 0 (ideal second accumulator-ish)", dict(fold="all8", tF=0)),
    ]:
        print(f"{simulate(**kw):8.0f} cycles/tile  {name}")
    print("-- round 2 model")
    for kw in [dict(), dict(tF=0), dict(tL=0), dict(tF=0, tL=0), dict(tF=1650), dict(tF=2300), dict(tC=350, tF=1650)]:
        print(f"{simulate_r2('cur', **kw):8.0f} cycles/tile as shipped   {simulate_r2('ring', **kw):8.0f} with the operand ring   {kw}")
