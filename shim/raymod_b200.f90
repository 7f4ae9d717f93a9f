!> Drop-in `module raymod` for Fortran callers of the reference ray tracer.
!>
!> The reference's sampler reaches the forward model through
!>     USE RAYMOD
!>     CALL TraceRays(vels, thickness, obj%k-1, src_offset, src_depth, NSRC, DpredRT, -1)
!> (ray_tracing_sampling/loglhood.f90:42,135,144; definition subroutineR-quiet.f90:467-482).
!> Compiling THIS file instead of subroutineR-quiet.f90 and linking libraytrace_b200.so keeps
!> loglhood.f90 and prjmh_temper_rf.f90 unchanged: TraceRays and dofullforwardproblem forward
!> to the C ABI (include/raytrace_b200.h), every argument by reference exactly as before.
!> The batched entries are exposed too, so a sampler that evaluates many proposals per step can
!> call the GPU once per step instead of once per proposal.
!>
!> NOTE: this image has no Fortran compiler, so this file is shipped as source and is not
!> compiled or tested here; the C ABI it binds is exercised by tests/ through ctypes with the
!> same all-by-reference convention (see INTEGRATION.md).
module raymod
   use, intrinsic :: iso_c_binding, only : c_double, c_int
   implicit none
   private
   public :: TraceRays, dofullforwardproblem, dff_batch, loglhood_batch, loglhood_batch_ar, loglhood_batch_voro

   interface
      ! void tracerays_(vels, depths, NLayers, src_offset, src_depth, NSrc, timeP, keep_delta)
      subroutine c_tracerays(vels, depths, NLayers, src_offset, src_depth, NSrc, timeP, keep_delta) &
            bind(C, name="tracerays_")
         import :: c_double, c_int
         real(c_double), intent(in)  :: vels(*), depths(*), src_offset(*), src_depth(*)
         integer(c_int), intent(in)  :: NLayers, NSrc, keep_delta
         real(c_double), intent(out) :: timeP(*)
      end subroutine c_tracerays

      ! int dff_batch(vels, depths, nlayers, B, ldv, ldz, src_offset, src_depth, NSrc,
      !               timeP, tobs, sigma, logL, p_out)   -- absent outputs: pass c_null_ptr via the
      ! type(c_ptr) overloads below if needed; this interface is for callers that want them all.
      function dff_batch(vels, depths, nlayers, B, ldv, ldz, src_offset, src_depth, NSrc, &
                         timeP, tobs, sigma, logL, p_out) bind(C, name="dff_batch") result(status)
         import :: c_double, c_int
         real(c_double), intent(in)  :: vels(*), depths(*), src_offset(*), src_depth(*), tobs(*), sigma(*)
         integer(c_int), intent(in)  :: nlayers(*), B, ldv, ldz, NSrc
         real(c_double), intent(out) :: timeP(*), logL(*), p_out(*)
         integer(c_int) :: status
      end function dff_batch

      ! int loglhood_batch(k, vp, ziface, B, ldv, ldz, src_offset, src_depth, NSrc, tobs, sigma,
      !                    logL, tpred)
      function loglhood_batch(k, vp, ziface, B, ldv, ldz, src_offset, src_depth, NSrc, &
                              tobs, sigma, logL, tpred) bind(C, name="loglhood_batch") result(status)
         import :: c_double, c_int
         integer(c_int), intent(in)  :: k(*), B, ldv, ldz, NSrc
         real(c_double), intent(in)  :: vp(*), ziface(*), src_offset(*), src_depth(*), tobs(*), sigma(*)
         real(c_double), intent(out) :: logL(*), tpred(*)
         integer(c_int) :: status
      end function loglhood_batch

      ! int loglhood_batch_ar(k, vp, ziface, B, ldv, ldz, src_offset, src_depth, NSrc, tobs, sigma,
      !                       idxar, arpar, armx, logL, tpred)      -- IAR = 1 (loglhood.f90:171-182)
      function loglhood_batch_ar(k, vp, ziface, B, ldv, ldz, src_offset, src_depth, NSrc, &
                                 tobs, sigma, idxar, arpar, armx, logL, tpred) &
            bind(C, name="loglhood_batch_ar") result(status)
         import :: c_double, c_int
         integer(c_int), intent(in)  :: k(*), B, ldv, ldz, NSrc, idxar(*)
         real(c_double), intent(in)  :: vp(*), ziface(*), src_offset(*), src_depth(*), tobs(*), sigma(*)
         real(c_double), intent(in)  :: arpar(*), armx
         real(c_double), intent(out) :: logL(*), tpred(*)
         integer(c_int) :: status
      end function loglhood_batch_ar

      ! int loglhood_batch_voro(k, voro, B, ldk, src_offset, src_depth, NSrc, tobs, sigma, logL,
      !                         tpred, voro_sorted)   -- INTERPLAYER_novar + LOGLHOOD on voro(ldk,2,B)
      function loglhood_batch_voro(k, voro, B, ldk, src_offset, src_depth, NSrc, tobs, sigma, &
                                   logL, tpred, voro_sorted) &
            bind(C, name="loglhood_batch_voro") result(status)
         import :: c_double, c_int
         integer(c_int), intent(in)  :: k(*), B, ldk, NSrc
         real(c_double), intent(in)  :: voro(*), src_offset(*), src_depth(*), tobs(*), sigma(*)
         real(c_double), intent(out) :: logL(*), tpred(*), voro_sorted(*)
         integer(c_int) :: status
      end function loglhood_batch_voro
   end interface

contains

   !> Same explicit-shape signature as subroutineR-quiet.f90:467-482.
   subroutine TraceRays(vels, depths, NLayers, src_offset, src_depth, NSrc, timeP, keep_delta)
      integer, intent(in) :: NLayers, NSrc
      integer, intent(in) :: keep_delta
      double precision, dimension(NSrc) :: src_offset, src_depth
      double precision, dimension(NSrc), intent(out) :: timeP
      double precision, dimension(NLayers+1), intent(in) :: vels
      double precision, dimension(NLayers), intent(in) :: depths
      call c_tracerays(vels, depths, int(NLayers, c_int), src_offset, src_depth, int(NSrc, c_int), &
                       timeP, int(keep_delta, c_int))
   end subroutine TraceRays

   !> Same signature as subroutineR-quiet.f90:408-425.  The C symbol dff_ itself is exported by
   !> libraytrace_b200.so (that is what R's .Fortran("dff", ...) resolves), so this procedure
   !> carries no bind(C) name of its own.
   subroutine dofullforwardproblem(vels, depths, NLayers, src_offset, src_depth, NSrc, timeP, keep_delta)
      integer(c_int), intent(in) :: NLayers, NSrc
      integer(c_int), intent(in) :: keep_delta
      real(c_double), dimension(NSrc) :: src_offset, src_depth
      real(c_double), dimension(NSrc), intent(out) :: timeP
      real(c_double), dimension(NLayers+1), intent(in) :: vels
      real(c_double), dimension(NLayers), intent(in) :: depths
      call c_tracerays(vels, depths, NLayers, src_offset, src_depth, NSrc, timeP, keep_delta)
   end subroutine dofullforwardproblem

end module raymod
