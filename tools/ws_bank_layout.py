"""Search for a bank-conflict-free placement of the consumer operands inside a tile-buffer row of grad_kernel_ws
(csrc/grad_kernel_ws.cuh: WS_A0..WS_A4, WS_D1A/B, WS_D2A/B).  Model: an LDS.128 of a warp is served in four phases of eight
lanes; within a phase, distinct 16-byte addresses that fall into the same bank group (address / 16 mod 8) serialise.  A consumer
lane (a-tile, delta half, row group) issues two a-tile loads and three delta-half loads per step; the cost of a layout is the
number of wavefronts of those five loads (20 = conflict-free).  python tools/ws_bank_layout.py"""
import itertools, random
def wavefronts(addr_units):  # list of 32 unit addresses for LDS.128; 4 phases of 8 lanes; each phase: #distinct addresses per bank max
    tot=0
    for ph in range(4):
        lanes=addr_units[8*ph:8*ph+8]
        banks={}
        for a in set(lanes):
            banks.setdefault(a%8,set()).add(a)
        tot+=max(len(v) for v in banks.values())
    return tot
def cost(S, aoff, doff, lanemap):
    # aoff[atile] unit offset (2 units), doff[(sel,half)] unit offset (3 units) sel 0=d1 (atile<2) 1=d2
    tot=0
    for k in range(2):
        addrs=[]
        for lane in range(32):
            combo,grp=lanemap(lane)
            atile,half=combo%5,combo//5
            addrs.append(grp*S+aoff[atile]+k)
        tot+=wavefronts(addrs)
    for k in range(3):
        addrs=[]
        for lane in range(32):
            combo,grp=lanemap(lane)
            atile,half=combo%5,combo//5
            addrs.append(grp*S+doff[(0 if atile<2 else 1,half)]+k)
        tot+=wavefronts(addrs)
    return tot
def lm_cur(lane):
    if lane>=30: return (0,0)
    return (lane%10, lane//10)
def lm_grpfast(lane):
    if lane>=30: return (0,0)
    return (lane//3, lane%3)


def header_layout(path):
    """(a-tile unit offsets, delta-half unit offsets) of the constants WS_A0..WS_A4, WS_D1A/B, WS_D2A/B in grad_kernel_ws.cuh"""
    import re
    txt = open(path).read()
    val = lambda name: int(re.search(r"\b" + name + r" = (\d+)", txt).group(1))
    for n in ("WS_A0", "WS_A1", "WS_A2", "WS_A3", "WS_A4", "WS_D1A", "WS_D1B", "WS_D2A", "WS_D2B"):
        assert val(n) % 4 == 0, n
    aoff = [val(f"WS_A{i}") // 4 for i in range(5)]
    doff = {(0, 0): val("WS_D1A") // 4, (0, 1): val("WS_D1B") // 4, (1, 0): val("WS_D2A") // 4, (1, 1): val("WS_D2B") // 4}
    return val("WS_ROWF") // 4, aoff, doff


if __name__ == "__main__":
    cur_a=[0,2,4,6,8]; cur_d={(0,0):10,(0,1):13,(1,0):16,(1,1):19}
    print("current", cost(23,cur_a,cur_d,lm_cur), "ideal", 5*4)
    best=None
    for lmname,lm in (("cur",lm_cur),("grpfast",lm_grpfast)):
      for S in (23,25):
        # place 9 regions + pads in S units: order permutations of regions with pad positions
        regs=[('a',0,2),('a',1,2),('a',2,2),('a',3,2),('a',4,2),('d',(0,0),3),('d',(0,1),3),('d',(1,0),3),('d',(1,1),3)]
        npad=S-22
        random.seed(1)
        for trial in range(200000):
            order=regs[:]; random.shuffle(order)
            # insert pads at random gaps
            gaps=[0]*(len(order)+1)
            for _ in range(npad): gaps[random.randrange(len(order)+1)]+=1
            pos=0; aoff=[0]*5; doff={}
            for i,r in enumerate(order):
                pos+=gaps[i]
                if r[0]=='a': aoff[r[1]]=pos
                else: doff[r[1]]=pos
                pos+=r[2]
            c=cost(S,aoff,doff,lm)
            if best is None or c<best[0]:
                best=(c,lmname,S,aoff[:],dict(doff)); print(best)
                if c==20: break
