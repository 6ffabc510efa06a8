// GpuFullRaytracer.cs — drop-in for RaytracerCore.Raytracing.FullRaytracer (Raytracing/FullRaytracer.cs) that renders
// on one B200 through librtcore_b200.so. Same public members as the reference class (ctor, Start, Stop, Pause, Resume,
// QueueUpdate, IsRunning/IsPaused/IsStopping, GetSampleSet, GetBitmap, Scene, Exposure), so MainWindow.cs:197 only
// changes the constructed type.
//
// NOT BUILT HERE (no .NET SDK in this repository's image). The C++ mirror that IS built and tested is
// raytracercore_b200/host/full_raytracer.cpp; this file is the same logic in the reference's own language.
//
// The reference keeps the fields this shim needs non-public (Triangle.Vert0/Edge0to1/Edge0to2/Normal/Mirror,
// Sphere.MatrixTo*, Plane.Normal/OriginDistance, AABB.Minimum/Maximum, Camera.look/side/w2/h2, FrustumCamera.tanFOV*,
// OrthoCamera.hMult/vMult); they are read by reflection below. Adding `internal` accessors would be the cleaner patch.
using System;
using System.Drawing;
using System.Drawing.Imaging;
using System.Reflection;
using System.Threading;
using RaytracerCore.Raytracing.Acceleration;
using RaytracerCore.Raytracing.Cameras;
using RaytracerCore.Raytracing.Primitives;
using RaytracerCore.Vectors;

namespace RaytracerCore.Raytracing.Gpu
{
	public unsafe class GpuFullRaytracer
	{
		public Scene Scene;
		public double Exposure = 1;

		readonly Action<GpuFullRaytracer, string, double, Bitmap> UpdateStatusCallback;
		readonly IntPtr Ctx;
		readonly object CtxLock = new object();   // one caller per handle (rtcore_b200.h)
		volatile bool Stopping, Running, Paused;
		readonly ManualResetEventSlim PauseWaiter = new ManualResetEventSlim(true);
		readonly uint SamplesPerPass;
		bool HaveImage;

		public GpuFullRaytracer(Scene scene, int device, Action<GpuFullRaytracer, string, double, Bitmap> updateStatus, uint samplesPerPass = 4)
		{
			Scene = scene;
			UpdateStatusCallback = updateStatus;
			SamplesPerPass = samplesPerPass;
			RtcoreNative.Check(IntPtr.Zero, RtcoreNative.rtc_create(device, RtcoreNative.F32, out Ctx));
		}

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
				if (p is Sphere s && Field<bool>(s, "Transformed")) nx++;
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
							// Triangle.HasNormals (trinormal): add a row holding Vert0/1/2.Normal and FlagVNormals
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
			// The reference's own accelerator can be handed over node by node (Left/Right/IsLeaf/LeafID/Volume) with
			// rtc_upload_bvh; for scenes beyond a few thousand primitives the library's builder is the practical choice.
			RtcoreNative.Check(Ctx, RtcoreNative.rtc_build_bvh(Ctx));
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

		/// <summary>Start the raytracer (blocking, like FullRaytracer.Start, FullRaytracer.cs:243).</summary>
		public void Start()
		{
			while (Running) ;
			Stopping = false; Running = true;
			UpdateStatusCallback?.Invoke(this, "Preparing scene...", 0, GetBitmap());
			lock (CtxLock)
			{
				UploadScene();
				UploadCameraAndParams((ulong)DateTime.Now.Ticks);
				RtcoreNative.Check(Ctx, RtcoreNative.rtc_clear_accum(Ctx));
				HaveImage = true;
			}
			UpdateStatusCallback?.Invoke(this, "Beginning render...", 0, GetBitmap());
			var watch = new System.Diagnostics.Stopwatch();
			uint done = 0; ulong passes = 0; TimeSpan total = TimeSpan.Zero;
			while (!Stopping)
			{
				PauseWaiter.Wait();   // workers park at a pass boundary (FullRaytracer.cs:223-224)
				if (Stopping) break;
				watch.Restart();
				lock (CtxLock)
				{
					RtcoreNative.Check(Ctx, RtcoreNative.rtc_render(Ctx, 0, 0, Scene.Width, Scene.Height, done, SamplesPerPass));
					RtcoreNative.Check(Ctx, RtcoreNative.rtc_sync(Ctx));
				}
				total += watch.Elapsed; done += SamplesPerPass; passes++;
				double perPixel = done, samplesPerSecond = perPixel / total.TotalSeconds, progress = perPixel / (perPixel + 1000);
				UpdateStatusCallback?.Invoke(this, $"Tiles: {passes:N0} Elapsed: {Util.FormatTimeSpan(total)} {perPixel:N2}/px {samplesPerSecond:N3}/px/sec", progress, GetBitmap());
			}
			Running = false;
		}

		public bool IsRunning => Running;
		public bool IsPaused => Paused;
		public bool IsStopping => Stopping;
		public void Pause() { Paused = true; PauseWaiter.Reset(); }
		public void Resume() { Paused = false; PauseWaiter.Set(); }
		public void Stop() { Stopping = true; Resume(); }
		public void QueueUpdate() { }   // status is pushed once per pass; nothing to wake

		public SampleSet GetSampleSet(int x, int y)
		{
			if (!HaveImage) return new SampleSet();
			x = Util.Clamp(x, 0, Scene.Width - 1); y = Util.Clamp(y, 0, Scene.Height - 1);
			int n = Scene.Width * Scene.Height;
			double[] rgb = new double[n * 3]; uint[] s = new uint[n], m = new uint[n];
			lock (CtxLock)
				fixed (double* r = rgb) fixed (uint* ps = s, pm = m)
					RtcoreNative.Check(Ctx, RtcoreNative.rtc_read_accum(Ctx, r, ps, pm));
			int i = y * Scene.Width + x;
			return new SampleSet(new DoubleColor(rgb[i * 3], rgb[i * 3 + 1], rgb[i * 3 + 2]), s[i], m[i]);
		}

		/// <summary>Convert the sample data to an output image (FullRaytracer.GetBitmap, FullRaytracer.cs:179-205) on the device.</summary>
		public Bitmap GetBitmap()
		{
			if (!HaveImage) return null;
			Bitmap bitmap = new Bitmap(Scene.Width, Scene.Height);
			BitmapData data = bitmap.LockBits(new Rectangle(0, 0, bitmap.Width, bitmap.Height), ImageLockMode.WriteOnly, PixelFormat.Format32bppArgb);
			double* back = stackalloc double[3];
			back[0] = Scene.BackgroundRGB.R; back[1] = Scene.BackgroundRGB.G; back[2] = Scene.BackgroundRGB.B;
			lock (CtxLock)
				RtcoreNative.Check(Ctx, RtcoreNative.rtc_tonemap_argb(Ctx, Exposure, back, Scene.BackgroundAlpha, (int*)data.Scan0.ToPointer()));
			bitmap.UnlockBits(data);
			return bitmap;
		}
	}
}
