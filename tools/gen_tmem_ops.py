"""Regenerates the tcgen05.ld / tcgen05.st operand lists of csrc/tmem_ops.cuh (32x32b shapes with 8, 16, 32 and 64
registers per thread); prints the helper functions to stdout.  The hand-written part of the header (alloc, fences,
tmem_load<N> / tmem_store<N>) is not generated."""
for n in (8, 16, 32, 64):
    regs_out = ", ".join(f"%{i}" for i in range(n))
    outs = ", ".join(f'"=r"(r[{i}])' for i in range(n))
    print(f'''
__device__ __forceinline__ void tmem_ld{n}(uint32_t taddr, uint32_t* r) {{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x{n}.b32 {{{regs_out}}}, [%{n}];"
                 : {outs}
                 : "r"(taddr));
}}''')
    regs_in = ", ".join(f"%{i + 1}" for i in range(n))
    ins = ", ".join(f'"r"(r[{i}])' for i in range(n))
    print(f'''
__device__ __forceinline__ void tmem_st{n}(uint32_t taddr, const uint32_t* r) {{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x{n}.b32 [%0], {{{regs_in}}};"
                 :: "r"(taddr), {ins} : "memory");
}}''')
