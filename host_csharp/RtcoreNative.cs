// RtcoreNative.cs — P/Invoke binding of librtcore_b200.so (include/rtcore_b200.h) for the reference's .NET host.
//
// NOT BUILT IN THIS REPOSITORY'S IMAGE: there is no .NET SDK here and the reference targets WinForms
// (RaytracerCore/RaytracerCore.csproj:1-11). The file is the binding a maintainer of Zaggy1024/RaytracerCore would
// add next to Raytracing/FullRaytracer.cs; every entry point mirrors the C header one to one.
using System;
using System.Runtime.InteropServices;

namespace RaytracerCore.Raytracing.Gpu
{
	[StructLayout(LayoutKind.Sequential)]
	public unsafe struct RtcSceneDesc
	{
		public int NPrims, NXforms;
		public byte* Kind, Flags;
		public double* Geom;
		public int* Xform;
		public double* Xforms;
		public double* Material;
	}

	[StructLayout(LayoutKind.Sequential)]
	public unsafe struct RtcBvhNode
	{
		public fixed double BMin[3];
		public fixed double BMax[3];
		public int Left, Right, Prim, Pad;
	}

	[StructLayout(LayoutKind.Sequential)]
	public unsafe struct RtcCamera
	{
		public int Kind, Pad;
		public fixed double Position[3];
		public fixed double Look[3];
		public fixed double Side[3];
		public fixed double Up[3];
		public double W2, H2, TanFovX2, TanFovY2, HMult, VMult, ImagePlane, DofAmount, FocalLength;
	}

	[StructLayout(LayoutKind.Sequential)]
	public unsafe struct RtcParams
	{
		public int Width, Height, Recursion, DebugGeom;
		public fixed double Ambient[3];
		public double AirIor;
		public ulong Seed;
	}

	[StructLayout(LayoutKind.Sequential)]
	public unsafe struct RtcRay { public fixed double Origin[3]; public fixed double Dir[3]; }

	[StructLayout(LayoutKind.Sequential)]
	public unsafe struct RtcHit
	{
		public int Prim, Inside;
		public double T;
		public fixed double Position[3];
		public fixed double Normal[3];
	}

	[StructLayout(LayoutKind.Sequential)]
	public struct RtcDebugRay { public RtcHit Hit; public int Type, Pad; public double FresnelRatio; }

	/// <summary>rtc_stats: RTC_K_COUNT = 5 kernel families (raygen, trace, shade, compact, accumulate).</summary>
	[StructLayout(LayoutKind.Sequential)]
	public unsafe struct RtcStats
	{
		public ulong Paths, Rays;
		public fixed ulong Launches[5];
		public fixed double Ms[5];
		public ulong NodesVisited, PrimsTested, NodeSteps, LeafSteps;
	}

	/// <summary>rtc_prepare_stats: phases of Scene.Prepare on the device (rtc_prepare_device).</summary>
	[StructLayout(LayoutKind.Sequential)]
	public struct RtcPrepareStats
	{
		public double BoxesMs, BuildMs, FlattenMs, TotalMs;
		public int BuildLevels, WideDepth, NWideNodes, NBounded;
	}

	public static unsafe class RtcoreNative
	{
		const string Lib = "rtcore_b200";   // librtcore_b200.so / rtcore_b200.dll

		public const int F32 = 0, F64 = 1;
		public const int BuilderSah = 0, BuilderPloc = 1;
		public const int KindTriangle = 0, KindSphere = 1, KindPlane = 2;
		public const int FlagMirror = 1, FlagTwoSided = 2, FlagInvert = 4, FlagTransformed = 8, FlagVNormals = 16;
		public const int OverlayPrimitives = 0, OverlayBoundingVolumes = 1;
		public const int OptKernelTiming = 1, OptCounters = 2, OptMaxPaths = 3;

		[DllImport(Lib)] public static extern int rtc_abi_version();
		[DllImport(Lib)] public static extern int rtc_device_count();
		[DllImport(Lib)] public static extern int rtc_create(int device, int precision, out IntPtr ctx);
		[DllImport(Lib)] public static extern void rtc_destroy(IntPtr ctx);
		[DllImport(Lib)] public static extern IntPtr rtc_last_error(IntPtr ctx);
		[DllImport(Lib)] public static extern int rtc_set_option(IntPtr ctx, int option, long value);
		[DllImport(Lib)] public static extern int rtc_set_stream(IntPtr ctx, IntPtr cudaStream);
		[DllImport(Lib)] public static extern int rtc_upload_scene(IntPtr ctx, RtcSceneDesc* scene);
		[DllImport(Lib)] public static extern int rtc_upload_bvh(IntPtr ctx, int nNodes, RtcBvhNode* nodes, int root);
		[DllImport(Lib)] public static extern int rtc_build_bvh(IntPtr ctx);
		[DllImport(Lib)] public static extern int rtc_bake(IntPtr ctx, out IntPtr baked);
		[DllImport(Lib)] public static extern int rtc_upload_baked(IntPtr ctx, IntPtr baked);
		[DllImport(Lib)] public static extern long rtc_baked_bytes(IntPtr baked);
		[DllImport(Lib)] public static extern void rtc_baked_free(IntPtr baked);
		[DllImport(Lib)] public static extern int rtc_baked_segment(IntPtr baked, int segment, out IntPtr data, out long bytes);
		[DllImport(Lib)] public static extern int rtc_prepare_device(IntPtr ctx, int builder, int radius, RtcPrepareStats* stats);
		[DllImport(Lib)] public static extern int rtc_get_bvh_size(IntPtr ctx, out int nNodes, out int root);
		[DllImport(Lib)] public static extern int rtc_get_bvh(IntPtr ctx, int capacity, RtcBvhNode* nodes);
		[DllImport(Lib)] public static extern int rtc_set_camera(IntPtr ctx, RtcCamera* camera);
		[DllImport(Lib)] public static extern int rtc_set_params(IntPtr ctx, RtcParams* p);
		[DllImport(Lib)] public static extern int rtc_trace_closest(IntPtr ctx, long n, RtcRay* rays, RtcHit* skip, RtcHit* hits);
		[DllImport(Lib)] public static extern int rtc_camera_rays(IntPtr ctx, long n, int* xy, uint* sample, RtcRay* rays);
		[DllImport(Lib)] public static extern int rtc_build_bvh_device(IntPtr ctx, int radius, out int rounds);
		[DllImport(Lib)] public static extern int rtc_debug_create_horizon(IntPtr ctx, long n, double* poleZTheta, double* xyz);
		[DllImport(Lib)] public static extern int rtc_render(IntPtr ctx, int x0, int y0, int x1, int y1, uint firstSample, uint nSamples);
		[DllImport(Lib)] public static extern int rtc_render_read(IntPtr ctx, uint firstSample, uint nSamples, double* rgbSum, uint* samples, uint* misses);
		[DllImport(Lib)] public static extern int rtc_sync(IntPtr ctx);
		[DllImport(Lib)] public static extern int rtc_clear_accum(IntPtr ctx);
		[DllImport(Lib)] public static extern int rtc_read_accum(IntPtr ctx, double* rgbSum, uint* samples, uint* misses);
		[DllImport(Lib)] public static extern int rtc_write_accum(IntPtr ctx, double* rgbSum, uint* samples, uint* misses);
		[DllImport(Lib)] public static extern int rtc_accum_device_ptrs(IntPtr ctx, out IntPtr rgbSum, out IntPtr samples, out IntPtr misses);
		[DllImport(Lib)] public static extern int rtc_read_pixel(IntPtr ctx, int x, int y, double* rgbSum, out uint samples, out uint misses);
		[DllImport(Lib)] public static extern int rtc_tonemap_argb(IntPtr ctx, double exposure, double* backRgb, double backA, uint* argb);
		[DllImport(Lib)] public static extern int rtc_debug_trace(IntPtr ctx, int x, int y, uint sample, int capacity, RtcDebugRay* rays, out int n);
		[DllImport(Lib)] public static extern int rtc_debug_raycast(IntPtr ctx, int mode, int* ids);
		[DllImport(Lib)] public static extern int rtc_debug_raycast_selection(IntPtr ctx, int nSel, int* primIds, int* ids);
		[DllImport(Lib)] public static extern int rtc_render_samples(IntPtr ctx, uint sample, double* rgb);
		[DllImport(Lib)] public static extern int rtc_get_stats(IntPtr ctx, RtcStats* stats);
		[DllImport(Lib)] public static extern int rtc_reset_stats(IntPtr ctx);
		[DllImport(Lib)] public static extern int rtc_comm_unique_id(byte* id128);
		[DllImport(Lib)] public static extern int rtc_comm_init(IntPtr ctx, int nRanks, int rank, byte* id128);
		[DllImport(Lib)] public static extern int rtc_reduce_accum(IntPtr ctx, int root);
		[DllImport(Lib)] public static extern int rtc_bcast_scene(IntPtr ctx, int root);
		[DllImport(Lib)] public static extern int rtc_comm_destroy(IntPtr ctx);

		public static void Check(IntPtr ctx, int rc)
		{
			if (rc != 0)
				throw new InvalidOperationException($"rtcore_b200 error {rc}: {Marshal.PtrToStringAnsi(rtc_last_error(ctx))}");
		}
	}
}
