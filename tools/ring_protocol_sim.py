"""Discrete simulation of the shared-memory ring protocol of gg_layer_bulk.cuh (producer + 8 consumer warps, full /
empty mbarriers with phases) under random schedules and random chunk ownership: asserts that no warp ever reads a
stage that holds another chunk, that no arrival falls into the wrong phase of an empty barrier, and that the
system never deadlocks.  The first version of the kernel (release without waiting for the chunk to land) fails
the second assertion here — and hung on the GPU.  Run: python tools/ring_protocol_sim.py"""
import random
def sim(seed, S=6, W=8, nblocks=5):
    rng=random.Random(seed)
    # per block: number of chunks, and for each warp the sorted list of chunks it has rows in
    blocks=[]
    for b in range(nblocks):
        n=rng.randint(0,14)
        own=[sorted(set(rng.randrange(n) for _ in range(rng.randint(0,4)))) if n else [] for w in range(W)]
        blocks.append((n,own))
    # barrier state
    full_phase=[0]*S; full_done=[0]*S     # full: completed phases count
    empty_cnt=[0]*S; empty_done=[0]*S
    loaded=set()  # sequence numbers loaded
    # producer state
    pq=0; seq=[]  # list of (block, chunk) in order
    for b,(n,own) in enumerate(blocks):
        for c in range(n): seq.append((b,c))
    total=len(seq)
    # warp programs: list of actions: ('wait_full', q) / ('arrive', q)
    progs=[]
    for w in range(W):
        prog=[]; qbase=0
        for b,(n,own) in enumerate(blocks):
            released=0; ready=0
            def release_to(c):
                nonlocal released, ready
                if c>released:
                    for k in range(released,c):
                        if k>=ready: prog.append(('wait',qbase+k))
                        prog.append(('arrive',qbase+k))
                    if c>ready: ready=c
                    released=c
            for c in own[w]:
                release_to(c)
                if c>=ready:
                    prog.append(('wait',qbase+c)); ready=c+1
                prog.append(('read',qbase+c))
            release_to(n)
            qbase+=n
        progs.append(prog)
    pc=[0]*W
    steps=0
    stage_owner=[None]*S
    while True:
        progress=False
        # producer
        if pq<total:
            s=pq%S
            ok = pq<S or empty_done[s] >= pq//S   # needs (pq//S) completions of empty[s]
            if ok:
                stage_owner[s]=pq; full_done[s]+=1; pq+=1; progress=True
        order=list(range(W)); rng.shuffle(order)
        for w in order:
            if pc[w]>=len(progs[w]): continue
            act,q=progs[w][pc[w]]; s=q%S
            if act=='wait':
                if full_done[s] >= q//S+1: pc[w]+=1; progress=True
            elif act=='read':
                assert stage_owner[s]==q, ("stale read",w,q,stage_owner[s])
                pc[w]+=1; progress=True
            else:
                # arrival must belong to phase q//S of empty[s]
                assert empty_done[s]==q//S, ("arrival in wrong phase",w,q,empty_done[s])
                empty_cnt[s]+=1
                if empty_cnt[s]==W: empty_cnt[s]=0; empty_done[s]+=1
                pc[w]+=1; progress=True
        if all(pc[w]>=len(progs[w]) for w in range(W)) and pq>=total: return True
        if not progress: raise RuntimeError(("deadlock",seed,pq,total,[ (pc[w],len(progs[w])) for w in range(W)]))
for seed in range(3000): sim(seed)
print("ok")
