# MotionPlanningB200.jl -- the binding a MotionPlanning.jl maintainer would add to switch the hot
# path to libmpb200.so (include/mpb200.h).  Julia-0.5-era syntax to match the reference; it cannot
# be executed in this repository's containers (no julia), so it is documentation-grade code: every
# ccall signature below is checked against the header by tests/test_abi.py.
#
# Seams used (all already present in the reference):
#   * helper_data_structures(V, dist) is overloaded per metric (geometric.jl:14, linearquadratic.jl:68)
#   * ImmutableNNC{T}(D::SparseMatrixCSC{T,Int}, r) is served by viewcol (nearneighbors.jl:23-28,128)
#   * SweptCollisionChecker subtypes implement is_free_state / is_free_motion and carry `count`
#     (collisioncheckers.jl:4-6, robots2D.jl:5-14, boxesND.jl:15-27)
module MotionPlanningB200

using MotionPlanning
import MotionPlanning: is_free_state, is_free_motion, is_free_path, helper_data_structures, inball!, inballF!, inballB!

const LIB = "libmpb200"
check(rc) = rc == 0 || error(unsafe_string(ccall((:mpb200_last_error, LIB), Cstring, ())))
init(device = 0) = check(ccall((:mpb200_init, LIB), Cint, (Cint,), device))

# ---- sample sets -------------------------------------------------------------------------------
type B200Samples
    h::Ptr{Void}
end
function B200Samples{S<:SVector}(V::Vector{S})
    h = Ref{Ptr{Void}}(C_NULL)
    M = statevec2mat(V)                      # zero-copy d x N view (primitivetypes.jl:21-23)
    check(ccall((:mpb200_samples_create, LIB), Cint, (Ptr{Float64}, Int64, Cint, Ref{Ptr{Void}}),
                M, size(M, 2), size(M, 1), h))
    s = B200Samples(h[])
    finalizer(s, x -> ccall((:mpb200_samples_destroy, LIB), Cint, (Ptr{Void},), x.h))
    s
end

# ---- Euclidean: the whole ImmutableNNC table in one call ------------------------------------------
"Precompute every r-ball on the GPU and install it as the sample set's ImmutableNNC (nearneighbors.jl:128)."
function precompute_inball!(NN::MetricNN, r::Float64)
    s = B200Samples(NN.V)
    t = Ref{Ptr{Void}}(C_NULL); nnz = Ref{Int64}(0)
    check(ccall((:mpb200_inball_build, LIB), Cint, (Ptr{Void}, Float64, Ref{Ptr{Void}}, Ref{Int64}), s.h, r, t, nnz))
    N = length(NN.V)
    colptr = Array(Int64, N + 1); rowval = Array(Int64, nnz[]); nzval = Array(Float64, nnz[])
    check(ccall((:mpb200_table_fetch, LIB), Cint, (Ptr{Void}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}),
                t[], colptr, rowval, nzval))
    D = SparseMatrixCSC(N, N, colptr, rowval, nzval)   # fields cross unchanged: 1-based Int64 / Float64
    NNi = MetricNN(NN.V, NN.dist, NN.init, ImmutableNNC(D, fill(r, N)), NN.DS, NN.US)
    NNi, s, t[]
end

# ---- collision checkers ----------------------------------------------------------------------------
type B200PointRobot2D <: SweptCollisionChecker
    cpu::PointRobot2D            # host-side shapes (constructors, inflate, plotting stay in Julia)
    h::Ptr{Void}
    count::Int
end
function pack(C::Compound2D)
    # flatten to the arrays of mpb200_obstacles2d_desc exactly like shapes2d.pack_obstacles:
    # Compound2D nodes -> gates (parent-before-child AABBs), Circle -> [c; r; xrange; yrange],
    # Polygon -> [xrange; yrange; points; normals; nextrema] (all precomputed by SAT2D.jl:12-51)
    error("see motionplanning.jl_b200/shapes2d.py: pack_obstacles -- a 30-line transliteration")
end
# state-level calls (fmt.jl:24,34,75 pass states): batches of one
function is_free_state(v::AbstractVector, CC::B200PointRobot2D, SS::StateSpace)
    out = Ref{UInt8}(0)
    check(ccall((:mpb200_states_free, LIB), Cint, (Ptr{Float64}, Int64, Cint, Ptr{Void}, Ptr{Void}, Ref{UInt8}),
                collect(v), 1, length(v), CC.h, space_desc(SS), out))
    out[] != 0
end
function is_free_motion(v::AbstractVector, w::AbstractVector, CC::B200PointRobot2D, SS::StateSpace)
    out = Ref{UInt8}(0)
    check(ccall((:mpb200_segments_free, LIB), Cint,
                (Ptr{Float64}, Ptr{Float64}, Int64, Cint, Ptr{Void}, Ptr{Void}, Ref{UInt8}),
                collect(v), collect(w), 1, length(v), CC.h, space_desc(SS), out))
    CC.count += 1
    out[] != 0
end

# ---- the drop-in change in fmtstar! ------------------------------------------------------------------
# Batched tables: F (point validity) and E (edge validity, aligned with the backward table: stored
# entry k of column x, row y  <=>  is_free_motion(V[y], V[x], CC, SS)).
#   F = BitVector(N); check(ccall((:mpb200_points_free, LIB), Cint, (Ptr{Void},Ptr{Void},Ptr{Void},Ptr{UInt64}),
#                                 s.h, CC.h, space_desc(SS), F.chunks))
#   E = BitVector(nnz); checks = Ref{Int64}(0)
#   check(ccall((:mpb200_edges_free, LIB), Cint, (Ptr{Void},Ptr{Void},Ptr{Void},Ptr{Void},Ptr{UInt64},Ref{Int64}),
#               s.h, t, CC.h, space_desc(SS), E.chunks, checks))
# and fmt.jl:72-75 becomes (one changed line; y_idx is already computed there):
#   neighborhood = nearB(P.V, x, r, H)          # still a viewcol of the ImmutableNNC
#   c_min, y_idx = findmin(C[nonzeroinds(neighborhood)] + nonzeros(neighborhood))
#   k = P.V.cache.D.colptr[x] - 1 + findnth(H[rowvals_of_column_x], y_idx)   # position of y_min in column x
#   if E[k]                                       # was: is_free_motion(P.V[y_min], P.V[x], P.CC, P.SS)

# ---- sample_free! on the device (sampling.jl:23-37) -------------------------------------------------------
# The uniform bulk of the sample set is drawn, filtered with is_free_state and compacted on the GPU; the handle
# is kept for the table builds, the host copy feeds P.V.  (Own Philox stream: deterministic in `seed`, not
# Julia's global RNG.)
function sample_free_b200{T}(SS::StateSpace{T}, CC::B200PointRobot2D, N::Int, seed::UInt64)
    d = dim(SS)
    M = Array(Float64, d, N)
    h = Ref{Ptr{Void}}(C_NULL); used = Ref{Int64}(0)
    check(ccall((:mpb200_sample_free, LIB), Cint,
                (Ptr{Void}, Ptr{Void}, Int64, UInt64, Int32, Ref{Ptr{Void}}, Ptr{Float64}, Ref{Int64}),
                CC.h, space_desc(SS), N, seed, Int32(1), h, M, used))   # 1 = Morton numbering
    s = B200Samples(h[])                              # adopt the handle
    finalizer(s, x -> ccall((:mpb200_samples_destroy, LIB), Cint, (Ptr{Void},), x.h))
    reinterpret(SVector{d,Float64}, M, (N,)), s, used[]
end

# ---- linear-quadratic (ControlNN) ---------------------------------------------------------------------
function helper_data_structures{S}(V::Vector{S}, M::LinearQuadratic, backend::Type{Val{:b200}})
    s = B200Samples(V)
    lq = Ref{Ptr{Void}}(C_NULL)
    b = M.bvp
    check(ccall((:mpb200_lq_create, LIB), Cint,
                (Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint, Cint, Ref{Ptr{Void}}),
                b.A, b.B, b.c, b.R, size(b.A, 1), size(b.B, 2), lq))
    tF = Ref{Ptr{Void}}(C_NULL); tB = Ref{Ptr{Void}}(C_NULL); nF = Ref{Int64}(0); nB = Ref{Int64}(0)
    check(ccall((:mpb200_lq_inball_build, LIB), Cint,
                (Ptr{Void}, Ptr{Void}, Float64, Ref{Ptr{Void}}, Ref{Ptr{Void}}, Ref{Int64}, Ref{Int64}),
                s.h, lq[], M.cmax, tF, tB, nF, nB))
    fetch(t, nnz) = begin
        N = length(V)
        cp = Array(Int64, N + 1); rv = Array(Int64, nnz); nz = Array(Float64, nnz)
        check(ccall((:mpb200_table_fetch, LIB), Cint, (Ptr{Void}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}), t, cp, rv, nz))
        SparseMatrixCSC(N, N, cp, rv, nz)
    end
    US = EmptyControlCache()
    BruteDistanceDS(fetch(tF[], nF[])), US, BruteDistanceDS(fetch(tB[], nB[])), US   # DSF, USF, DSB, USB (:73-76)
end

end # module
