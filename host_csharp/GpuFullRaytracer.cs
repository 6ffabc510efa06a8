// GpuFullRaytracer.cs — drop-in for RaytracerCore.Raytracing.FullRaytracer (Raytracing/FullRaytracer.cs) that renders
// on one B200 through librtcore_b200.so. It has the reference class's whole public surface, member for member:
//
//   reference (FullRaytracer.cs)                                   here
//   :35-36   Scene, Exposure                                       Scene, Exposure
//   :61,63   DebugPathtracer (Raytracer), DebugRaycaster           DebugPathtracer (GpuDebugPathtracer), DebugRaycaster (GpuDebugRaycaster)
//   :66      FullRaytracer(scene, threads, updateStatus, updateDebug)   same four arguments (threads is ignored; the GPU is RTCORE_DEVICE or 0)
//   :131     GetSampleSet(x, y)                                    rtc_read_pixel
//   :179     GetBitmap()                                           rtc_tonemap_argb (device tonemap, copy stream)
//   :243     Start()                                               blocking progressive loop over rtc_render
//   :375-416 IsRunning, Pause, QueueUpdate, QueueDebugUpdate, IsPaused, Resume, Stop, IsStopping
//
// so the UI switches over with one alias line per file that names the type (MainWindow.cs, Inspector/RayInspector.cs):
//     using FullRaytracer = RaytracerCore.Raytracing.Gpu.GpuFullRaytracer;
// `new FullRaytracer(scene, Environment.ProcessorCount, UpdateRenderedImage, UpdateDebugImage)` (MainWindow.cs:197),
// `CurrentRaytracer.QueueDebugUpdate()` (:357), `CurrentRaytracer.DebugRaycaster.SetMode(DebugRaycaster.DisplayMode.Primitives)`
// (:377-380), `.SetDisplayOnly(...)` / `.ClearDisplayOnly()` (:413,:422) and `Raytracer.DebugPathtracer.GetDebugTrace(X, Y)`
// under `lock (Raytracer.DebugPathtracer)` (RayInspector.cs:141-147) then compile unchanged.
//
// NOT BUILT HERE (no .NET SDK in this repository's image). The C++ mirror that IS built and tested is
// raytracercore_b200/host/full_raytracer.cpp; tests/test_csharp_binding.py checks this file's P/Invoke surface and struct
// layouts against include/rtcore_b200.h.
//
// The reference keeps the fields this shim needs non-public (Triangle.Vert0/Edge0to1/Edge0to2/Normal/Mirror,
// Sphere.MatrixTo*, Plane.Normal/OriginDistance, AABB.Minimum/Maximum, Camera.look/side/w2/h2, FrustumCamera.tanFOV*,
// OrthoCamera.hMult/vMult); they are read by reflection below. Adding `internal` accessors would be the cleaner patch.
using System;
using System.Collections.Generic;
using System.Drawing;
using System.Drawing.Imaging;
using System.Linq;
using System.Reflection;
using System.Threading;
using RaytracerCore.Raytracing.Acceleration;
using RaytracerCore.Raytracing.Cameras;
using RaytracerCore.Raytracing.Primitives;
using RaytracerCore.Vectors;

namespace RaytracerCore.Raytracing.Gpu
{
	/// <summary>Raytracer.GetDebugTrace(x, y) (Raytracer.cs:254-260, 289-292) over rtc_debug_trace. RayInspector locks this object.</summary>
	public unsafe class GpuDebugPathtracer
	{
		readonly GpuFullRaytracer Owner;
		uint NextSample = 0x40000000;   // a sample range of its own: inspector traces never replay a rendered sample

		internal GpuDebugPathtracer(GpuFullRaytracer owner) { Owner = owner; }

		public Raytracer.DebugRay[] GetDebugTrace(int x, int y)
		{
			Scene scene = Owner.Scene;
			int cap = scene.Recursion + 1;   // Raytracer.cs:257
			RtcDebugRay[] raw = new RtcDebugRay[cap];
			int n;
			lock (Owner.CtxLock)
			{
				if (!Owner.EnsureSceneOnDevice()) return new Raytracer.DebugRay[0];
				fixed (RtcDebugRay* p = raw)
					RtcoreNative.Check(Owner.Ctx, RtcoreNative.rtc_debug_trace(Owner.Ctx, x, y, NextSample++, cap, p, out n));
			}
			var trace = new Raytracer.DebugRay[n];
			IList<Primitive> prims = scene.Primitives;
			for (int i = 0; i < n; i++)
			{
				RtcDebugRay r = raw[i];
				var d = new Raytracer.DebugRay { Type = (Raytracer.BounceType)r.Type, FresnelRatio = r.FresnelRatio };
				if (r.Hit.Prim >= 0)
					d.Hit = new Hit(prims[r.Hit.Prim], new Vec4D(r.Hit.Position[0], r.Hit.Position[1], r.Hit.Position[2], 1), r.Hit.T,
						new Vec4D(r.Hit.Normal[0], r.Hit.Normal[1], r.Hit.Normal[2], 0), r.Hit.Inside != 0);
				trace[i] = d;
			}
			return trace;
		}
	}

	/// <summary>DebugRaycaster (Raytracing/DebugRaycaster.cs) with the per-pixel queries of the Primitives and BoundingVolumes
	/// modes (:194-212) answered by rtc_debug_raycast and the Selection mode over primitives (:174-192) by
	/// rtc_debug_raycast_selection; only a selected BVH node (one box test per pixel) stays on the reference's own CPU code. Derives from the reference class so that DisplayMode, ColorRotation and
	/// the intersector types are the reference's; the four public methods are re-declared because the originals are not virtual.</summary>
	public unsafe class GpuDebugRaycaster : DebugRaycaster
	{
		readonly GpuFullRaytracer Owner;
		DisplayMode GpuNextMode = DisplayMode.Primitives;   // DebugRaycaster.NextMode is private (:108)
		bool HaveSelection;
		int MaxBoxes = -1;                                   // DebugRaycaster.MaxBoundingBoxes (:111)

		internal GpuDebugRaycaster(GpuFullRaytracer owner, Scene scene) : base(scene) { Owner = owner; }

		public new void SetMode(DisplayMode mode)   // :118-128
		{
			if (mode == DisplayMode.Selection && !HaveSelection)
			{
				if (GpuNextMode == DisplayMode.Selection) GpuNextMode = DisplayMode.Primitives;
				return;
			}
			GpuNextMode = mode;
			base.SetMode(mode);
		}

		public new void ClearDisplayOnly() { HaveSelection = false; base.ClearDisplayOnly(); }   // :130-133

		int[] SelectedPrimitives;   // Primitive.IDs of a Primitive / IObject selection (null: none, or a BVH node was selected)

		public new bool SetDisplayOnly(object item)   // :140-165
		{
			HaveSelection = false;
			SelectedPrimitives = null;
			bool valid = base.SetDisplayOnly(item);   // builds the intersector list and switches the base to Selection
			HaveSelection = NextIntersectors != null;
			if (item is Primitive primitive) SelectedPrimitives = new[] { Math.Max(primitive.ID, 0) };
			else if (item is IObject parent) SelectedPrimitives = Scene.Primitives.Where(p => p.Parent == parent).Select(p => Math.Max(p.ID, 0)).ToArray();
			SetMode(DisplayMode.Selection);
			return valid;
		}

		public new Bitmap RenderDebug()   // :217-265
		{
			DisplayMode mode = GpuNextMode;
			int[] selection = SelectedPrimitives;
			// a selected BVH node is one box test per pixel: the reference's own loop; a selection of primitives (a whole object may be
			// a million triangles, each tested for every pixel by the reference, :174-192) goes to the device
			if (mode == DisplayMode.Selection && selection == null) return base.RenderDebug();
			int w = Scene.Width, h = Scene.Height;
			int[] ids = new int[w * h];
			lock (Owner.CtxLock)
			{
				if (!Owner.EnsureSceneOnDevice()) return null;
				fixed (int* p = ids)
				{
					if (mode == DisplayMode.Selection)
						fixed (int* sel = selection)
							RtcoreNative.Check(Owner.Ctx, RtcoreNative.rtc_debug_raycast_selection(Owner.Ctx, selection.Length, sel, p));
					else
						RtcoreNative.Check(Owner.Ctx, RtcoreNative.rtc_debug_raycast(Owner.Ctx,
							mode == DisplayMode.BoundingVolumes ? RtcoreNative.OverlayBoundingVolumes : RtcoreNative.OverlayPrimitives, p));
				}
			}
			if (mode == DisplayMode.BoundingVolumes)
				foreach (int c in ids)
					if (c > MaxBoxes) MaxBoxes = c;   // :207-208 (running maximum, as in the reference)
			Bitmap output = new Bitmap(w, h);
			BitmapData data = output.LockBits(new Rectangle(0, 0, w, h), ImageLockMode.WriteOnly, PixelFormat.Format32bppArgb);
			int* values = (int*)data.Scan0.ToPointer();
			for (int y = 0; y < h; y++)
				for (int x = 0; x < w; x++)
				{
					int v = ids[y * w + x];
					Color color;
					if (mode == DisplayMode.BoundingVolumes)
					{
						if (v == 0) color = Color.Transparent;   // :210-211
						else
						{
							color = Color.FromArgb(Math.Min(v, 255), 255, 255, 255);   // :213
							double a = Math.Sqrt(color.A / (double)MaxBoxes);       // :245-246
							color = Color.FromArgb((int)(a * 255), color);
						}
					}
					else
					{
						color = v < 0 ? Color.Transparent : ColorRotation[v % ColorRotation.Length];   // :196-199, :167-172
						color = Color.FromArgb(color.A / 2, color);                                    // :249
					}
					values[(y * data.Width) + x] = color.ToArgb();
				}
			output.UnlockBits(data);
			return output;
		}
	}

	public unsafe class GpuFullRaytracer
	{
		const int Interval = 100;   // FullRaytracer.cs:33

		public Scene Scene;
		public double Exposure = 1;

		public readonly GpuDebugPathtracer DebugPathtracer;   // FullRaytracer.cs:61
		public readonly GpuDebugRaycaster DebugRaycaster;     // FullRaytracer.cs:63

		readonly Action<GpuFullRaytracer, string, double, Bitmap> UpdateStatusCallback;
		readonly Action<GpuFullRaytracer, Bitmap> UpdateDebugCallback;
		internal readonly IntPtr Ctx;
		internal readonly object CtxLock = new object();   // one caller per handle (rtcore_b200.h)
		volatile bool Stopping, Running, Paused, DebugChanged;
		readonly EventWaitHandle IntervalWaiter = new EventWaitHandle(true, EventResetMode.AutoReset);   // :88
		readonly EventWaitHandle PauseWaiter = new EventWaitHandle(true, EventResetMode.ManualReset);    // :90
		readonly uint SamplesPerPass;
		bool HaveImage, SceneOnDevice;

		/// <summary>The reference's constructor (FullRaytracer.cs:66). `threads` is accepted for source compatibility; the device
		/// is taken from the RTCORE_DEVICE environment variable (default 0).</summary>
		public GpuFullRaytracer(Scene scene, int threads, Action<GpuFullRaytracer, string, double, Bitmap> updateStatus,
			Action<GpuFullRaytracer, Bitmap> updateDebug)
		{
			Scene = scene;
			UpdateStatusCallback = updateStatus;
			UpdateDebugCallback = updateDebug;
			int device = 0;
			int.TryParse(Environment.GetEnvironmentVariable("RTCORE_DEVICE"), out device);
			// samples per pass: RTCORE_SAMPLES_PER_PASS, else passes of about 8 Mi paths (what keeps both wavefronts of rtc_render busy
			// while a pass still ends, and Stop / Pause / the status line are served, every few tens of milliseconds)
			uint spp;
			if (!uint.TryParse(Environment.GetEnvironmentVariable("RTCORE_SAMPLES_PER_PASS"), out spp) || spp == 0)
				spp = (uint)Math.Min(64.0, Math.Max(1.0, Math.Ceiling((double)(8u << 20) / Math.Max(1.0, (double)scene.Width * scene.Height))));
			SamplesPerPass = spp;
			RtcoreNative.Check(IntPtr.Zero, RtcoreNative.rtc_create(device, RtcoreNative.F32, out Ctx));
			DebugPathtracer = new GpuDebugPathtracer(this);
			DebugRaycaster = new GpuDebugRaycaster(this, scene);
		}

		~GpuFullRaytracer() { RtcoreNative.rtc_destroy(Ctx); }

		static T Field<T>(object o, string name)
		{
			for (Type t = o.GetType(); t != null; t = t.BaseType)
			{
				FieldInfo f = t.GetField(name, BindingFlags.Instance | BindingFlags.NonPublic | BindingFlags.Public);
				if (f != null) return (T)f.GetValue(o);
			}
			throw new MissingFieldException(o.GetType().Name, name);
		}

		static void Put(double* dst, Vec4D v) { dst[0] = v.X; dst[1] = v.Y; dst[2] = v.Z; }

		static void PutMatrix(double* dst, Mat4x4D m)
		{
			double[] d = { m.D00, m.D01, m.D02, m.D03, m.D10, m.D11, m.D12, m.D13, m.D20, m.D21, m.D22, m.D23, m.D30, m.D31, m.D32, m.D33 };
			for (int i = 0; i < 16; i++) dst[i] = d[i];
		}

		// Scene.Primitives -> rtc_scene_desc (Primitive.ID order), Scene.Accelerator -> rtc_bvh_node[]
		void UploadScene()
		{
			var prims = Scene.Primitives;
			int n = prims.Count, nx = 0;
			foreach (Primitive p in prims)
				if ((p is Sphere s && Field<bool>(s, "Transformed")) || (p is Triangle t && Field<bool>(t, "HasNormals"))) nx++;
			byte[] kind = new byte[n], flags = new byte[n];
			double[] geom = new double[n * 12], material = new double[n * 14], xforms = new double[Math.Max(1, nx) * 48];
			int[] xform = new int[n];
			fixed (double* g = geom, m = material, x = xforms)
			{
				int xi = 0;
				for (int i = 0; i < n; i++)
				{
					Primitive p = prims[i];
					byte f = 0;
					if (p.TwoSided) f |= RtcoreNative.FlagTwoSided;
					if (p.Invert) f |= RtcoreNative.FlagInvert;
					xform[i] = -1;
					switch (p)
					{
						case Triangle t:
							kind[i] = RtcoreNative.KindTriangle;
							Put(g + i * 12, Field<Vertex>(t, "Vert0").Position);
							Put(g + i * 12 + 3, Field<Vec4D>(t, "Edge0to1"));
							Put(g + i * 12 + 6, Field<Vec4D>(t, "Edge0to2"));
							Put(g + i * 12 + 9, Field<Vec4D>(t, "Normal"));
							if (Field<bool>(t, "Mirror")) f |= RtcoreNative.FlagMirror;
							if (Field<bool>(t, "HasNormals"))   // trinormal (Triangle.cs:26,59-66): Vert0/1/2.Normal in the row
							{
								f |= RtcoreNative.FlagVNormals;
								xform[i] = xi;
								Put(x + xi * 48, Field<Vertex>(t, "Vert0").Normal);
								Put(x + xi * 48 + 3, Field<Vertex>(t, "Vert1").Normal);
								Put(x + xi * 48 + 6, Field<Vertex>(t, "Vert2").Normal);
								xi++;
							}
							break;
						case Sphere s:
							kind[i] = RtcoreNative.KindSphere;
							Put(g + i * 12, s.Center);
							g[i * 12 + 3] = s.Radius;
							g[i * 12 + 4] = s.Radius * s.Radius;
							if (Field<bool>(s, "Transformed"))
							{
								f |= RtcoreNative.FlagTransformed;
								xform[i] = xi;
								PutMatrix(x + xi * 48, Field<Mat4x4D>(s, "MatrixToWorld"));
								PutMatrix(x + xi * 48 + 16, Field<Mat4x4D>(s, "MatrixToObject"));
								PutMatrix(x + xi * 48 + 32, Field<Mat4x4D>(s, "MatrixToNormal"));
								xi++;
							}
							break;
						case Plane pl:
							kind[i] = RtcoreNative.KindPlane;
							Put(g + i * 12, Field<Vec4D>(pl, "Normal"));
							g[i * 12 + 3] = Field<double>(pl, "OriginDistance");
							break;
					}
					flags[i] = f;
					// raw backing fields: the IsReflective gating of Specular/Refraction (Primitive.cs:111-129) is applied by the library
					DoubleColor spec = Field<DoubleColor>(p, "_Specular"), refr = Field<DoubleColor>(p, "_Refraction");
					double[] mm = { p.Emission.R, p.Emission.G, p.Emission.B, p.Diffuse.R, p.Diffuse.G, p.Diffuse.B,
						spec.R, spec.G, spec.B, refr.R, refr.G, refr.B, p.RefractiveIndex, p.Shininess };
					for (int k = 0; k < 14; k++) m[i * 14 + k] = mm[k];
				}
				fixed (byte* k = kind, fl = flags)
				fixed (int* xf = xform)
				{
					RtcSceneDesc d = new RtcSceneDesc { NPrims = n, NXforms = nx, Kind = k, Flags = fl, Geom = g, Xform = xf, Xforms = x, Material = m };
					RtcoreNative.Check(Ctx, RtcoreNative.rtc_upload_scene(Ctx, &d));
				}
			}
			if (n <= ReferenceBuilderLimit)
			{
				// small scenes: Scene.Prepare (Scene.cs:39-49) builds the reference's own accelerator, which is handed over node by
				// node so that SceneInspector (SceneInspector.cs:226-265) shows exactly the tree being traced
				Scene.Prepare();
				var nodes = new List<RtcBvhNode>();
				int root = Flatten(Scene.Accelerator, nodes);
				RtcBvhNode[] arr = nodes.ToArray();
				fixed (RtcBvhNode* pn = arr)
					RtcoreNative.Check(Ctx, RtcoreNative.rtc_upload_bvh(Ctx, arr.Length, pn, root));
			}
			else
			{
				// the reference's agglomerative build is quadratic in practice (BVH.cs:50-191): the library takes over, tree and device
				// layout both made on the GPU (30 ms for a million triangles); SceneInspector reads the tree back with rtc_get_bvh
				RtcoreNative.Check(Ctx, RtcoreNative.rtc_prepare_device(Ctx, RtcoreNative.BuilderSah, 0, null));
			}
		}

		const int ReferenceBuilderLimit = 20000;

		// BVH<Primitive> (BVH.cs:239-254) -> rtc_bvh_node[], children before parents; returns the node's index
		static int Flatten(BVH<Primitive> node, List<RtcBvhNode> output)
		{
			RtcBvhNode n = new RtcBvhNode();
			Vec4D lo = Field<Vec4D>(node.Volume, "Minimum"), hi = Field<Vec4D>(node.Volume, "Maximum");
			n.BMin[0] = lo.X; n.BMin[1] = lo.Y; n.BMin[2] = lo.Z;
			n.BMax[0] = hi.X; n.BMax[1] = hi.Y; n.BMax[2] = hi.Z;
			if (node.IsLeaf) { n.Left = n.Right = -1; n.Prim = node.LeafID; }
			else { n.Left = Flatten(node.Left, output); n.Right = Flatten(node.Right, output); n.Prim = -1; }
			output.Add(n);
			return output.Count - 1;
		}

		void UploadCameraAndParams(ulong seed)
		{
			Camera cam = Scene.Camera;
			cam.InitRender(Scene.Width, Scene.Height);   // FullRaytracer.cs:269
			RtcCamera c = new RtcCamera();
			c.Kind = cam is OrthoCamera ? 1 : 0;
			Put(c.Position, cam.position); Put(c.Look, Field<Vec4D>(cam, "look")); Put(c.Side, Field<Vec4D>(cam, "side")); Put(c.Up, cam.up);
			c.W2 = Field<double>(cam, "w2"); c.H2 = Field<double>(cam, "h2");
			if (cam is FrustumCamera) { c.TanFovX2 = Field<double>(cam, "tanFOVX2"); c.TanFovY2 = Field<double>(cam, "tanFOVY2"); }
			else { c.HMult = Field<double>(cam, "hMult"); c.VMult = Field<double>(cam, "vMult"); }
			c.ImagePlane = cam.imagePlane; c.DofAmount = cam.dofAmount; c.FocalLength = cam.focalLength;
			RtcoreNative.Check(Ctx, RtcoreNative.rtc_set_camera(Ctx, &c));
			RtcParams p = new RtcParams { Width = Scene.Width, Height = Scene.Height, Recursion = Scene.Recursion, DebugGeom = Scene.DebugGeom ? 1 : 0,
				AirIor = Scene.AirRefractiveIndex, Seed = seed };
			p.Ambient[0] = Scene.AmbientRGB.R; p.Ambient[1] = Scene.AmbientRGB.G; p.Ambient[2] = Scene.AmbientRGB.B;
			RtcoreNative.Check(Ctx, RtcoreNative.rtc_set_params(Ctx, &p));
		}

		/// <summary>Scene + camera on the device (called with CtxLock held): the inspector and the overlay may query before Start().</summary>
		internal bool EnsureSceneOnDevice()
		{
			if (SceneOnDevice) return true;
			if (Scene == null || Scene.Primitives.Count == 0) return false;
			UploadScene();
			UploadCameraAndParams((ulong)DateTime.Now.Ticks);
			SceneOnDevice = true;
			return true;
		}

		void UpdateStatus(string status, double progress)   // :91-94
		{
			UpdateStatusCallback?.Invoke(this, status, progress, GetBitmap());
		}

		void UpdateDebug()   // :231-238
		{
			if (DebugChanged)
			{
				DebugChanged = false;
				UpdateDebugCallback?.Invoke(this, DebugRaycaster.RenderDebug());
			}
		}

		/// <summary>Start the raytracer (blocking, like FullRaytracer.Start, FullRaytracer.cs:243).</summary>
		public void Start()
		{
			while (Running) ;
			Stopping = false; Running = true;
			UpdateStatus("Preparing scene...", 0);
			lock (CtxLock)
			{
				SceneOnDevice = false;   // Start() always takes the scene as it is now (FullRaytracer.cs:253-269)
				EnsureSceneOnDevice();
				RtcoreNative.Check(Ctx, RtcoreNative.rtc_clear_accum(Ctx));
				HaveImage = true;
			}
			UpdateStatus("Beginning render...", 0);
			var watch = new System.Diagnostics.Stopwatch();
			var sleepTimer = new System.Diagnostics.Stopwatch();
			uint done = 0; ulong passes = 0; TimeSpan total = TimeSpan.Zero;
			while (!Stopping)
			{
				sleepTimer.Restart();
				watch.Restart();
				lock (CtxLock)
				{
					RtcoreNative.Check(Ctx, RtcoreNative.rtc_render(Ctx, 0, 0, Scene.Width, Scene.Height, done, SamplesPerPass));
					RtcoreNative.Check(Ctx, RtcoreNative.rtc_sync(Ctx));
				}
				total += watch.Elapsed; done += SamplesPerPass; passes++;
				double perPixel = done, samplesPerSecond = perPixel / total.TotalSeconds, progress = perPixel / (perPixel + 1000);   // :352-357
				UpdateStatus($"Tiles: {passes:N0} Elapsed: {Util.FormatTimeSpan(total)} {perPixel:N2}/px {samplesPerSecond:N3}/px/sec", progress);
				UpdateDebug();   // :363
				if (Paused)      // :365-371: workers park at a pass boundary
				{
					PauseWaiter.WaitOne();
					if (Paused) PauseWaiter.Reset();
				}
				// the reference sleeps out the rest of the 100 ms interval (:373) because its workers render meanwhile; here the
				// loop IS the renderer, so it only yields to a pending QueueUpdate
				IntervalWaiter.WaitOne(0);
			}
			Running = false;
		}

		public bool IsRunning => Running;                                   // :375
		public void Pause() { Paused = true; PauseWaiter.Reset(); }         // :377-381
		public void QueueUpdate()                                           // :383-393
		{
			IntervalWaiter.Set();
			if (Paused) PauseWaiter.Set();   // one more loop turn (status + debug overlay), then parked again
		}
		public void QueueDebugUpdate()                                      // :395-399
		{
			DebugChanged = true;
			QueueUpdate();
			if (!Running)   // the reference only refreshes the overlay from its render loop; without one, do it here
				ThreadPool.QueueUserWorkItem(_ => UpdateDebug());
		}
		public bool IsPaused => Paused;                                     // :401
		public void Resume() { Paused = false; PauseWaiter.Set(); }         // :403-407
		public void Stop() { Stopping = true; Resume(); }                   // :409-414
		public bool IsStopping => Stopping;                                 // :416

		/// <summary>GetSampleSet(x, y) (FullRaytracer.cs:131-146): one pixel of the accumulation planes.</summary>
		public SampleSet GetSampleSet(int x, int y)
		{
			if (!HaveImage) return new SampleSet();
			x = Util.Clamp(x, 0, Scene.Width - 1); y = Util.Clamp(y, 0, Scene.Height - 1);   // :137-138
			double* rgb = stackalloc double[3];
			uint s, m;
			lock (CtxLock)
				RtcoreNative.Check(Ctx, RtcoreNative.rtc_read_pixel(Ctx, x, y, rgb, out s, out m));
			return new SampleSet(new DoubleColor(rgb[0], rgb[1], rgb[2]), s, m);
		}

		/// <summary>Convert the sample data to an output image (FullRaytracer.GetBitmap, FullRaytracer.cs:179-205) on the device.</summary>
		public Bitmap GetBitmap()
		{
			if (!HaveImage) return null;   // :184-185
			Bitmap bitmap = new Bitmap(Scene.Width, Scene.Height);
			BitmapData data = bitmap.LockBits(new Rectangle(0, 0, bitmap.Width, bitmap.Height), ImageLockMode.WriteOnly, PixelFormat.Format32bppArgb);
			double* back = stackalloc double[3];
			back[0] = Scene.BackgroundRGB.R; back[1] = Scene.BackgroundRGB.G; back[2] = Scene.BackgroundRGB.B;
			lock (CtxLock)
				RtcoreNative.Check(Ctx, RtcoreNative.rtc_tonemap_argb(Ctx, Exposure, back, Scene.BackgroundAlpha, (uint*)data.Scan0.ToPointer()));
			bitmap.UnlockBits(data);
			return bitmap;
		}

		/// <summary>Counters behind the status line (FullRaytracer.cs:346-357) plus per-kernel device times.</summary>
		public RtcStats GetStats()
		{
			RtcStats st;
			lock (CtxLock)
				RtcoreNative.Check(Ctx, RtcoreNative.rtc_get_stats(Ctx, &st));
			return st;
		}
	}
}
